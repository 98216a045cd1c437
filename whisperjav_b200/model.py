"""B200-native Whisper model object: the drop-in for what ``whisper.load_model(name, device)``
returns in the reference (whisperjav/modules/whisper_pro_asr.py:182) -- same ``transcribe(audio,
**params)`` call and result dict (whisper_pro_asr.py:433; upstream whisper/transcribe.py).

Python here owns tensors, host<->device copies and the seek/threshold bookkeeping; every
arithmetic step (log-mel, encoder, cross-K/V, the whole greedy decode loop with its logit
filters) runs in hand-written sm_100a kernels behind the C-ABI of libwjb200.so.  There is no
CPU path: constructing the model without a CUDA device or without the library raises.

Batching: the reference decodes one <=30 s window per ``model.decode`` call; here any number of
independent clips advance through their seek loops in lock-step, ``max_batch`` windows per
device pass (windows are independent because every reference preset sets
``condition_on_previous_text=False``; prompts are still honoured per clip).
"""
from __future__ import annotations

import ctypes as C
import math
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from .hostlogic import beam_finalize_and_rank
from .synth import DIMS, Dims, synth_preset, synth_weights
from .weights import pack_weights

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
N_SAMPLES = 480000
N_FRAMES = 3000
# upstream whisper/tokenizer.py::LANGUAGES, in token-id order (<|en|> = sot + 1, ...); the last entry ("yue") exists only in the
# 51866-token vocabulary of large-v3 / turbo
LANGUAGES = ("en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it", "id", "hi", "fi", "vi",
             "he", "uk", "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur", "hr", "bg", "lt", "la", "mi", "ml", "cy", "sk",
             "te", "fa", "lv", "bn", "sr", "az", "sl", "kn", "et", "mk", "br", "eu", "is", "hy", "ne", "mn", "bs", "kk", "sq", "sw",
             "gl", "mr", "pa", "si", "km", "sn", "yo", "so", "af", "oc", "ka", "be", "tg", "sd", "gu", "am", "yi", "lo", "uz", "fo",
             "ht", "ps", "tk", "nn", "mt", "sa", "lb", "my", "bo", "tl", "mg", "as", "tt", "haw", "ln", "ha", "ba", "jw", "su", "yue")
# upstream tokenizer.py::TO_LANGUAGE_CODE aliases that reference configs use ("japanese" -> "ja", ...)
LANGUAGE_ALIASES = {"english": "en", "chinese": "zh", "german": "de", "spanish": "es", "russian": "ru", "korean": "ko", "french": "fr",
                    "japanese": "ja", "portuguese": "pt", "turkish": "tr", "polish": "pl", "dutch": "nl", "arabic": "ar", "swedish": "sv",
                    "italian": "it", "indonesian": "id", "hindi": "hi", "finnish": "fi", "vietnamese": "vi", "cantonese": "yue",
                    "mandarin": "zh", "burmese": "my", "castilian": "es", "flemish": "nl", "haitian": "ht", "moldavian": "ro",
                    "moldovan": "ro", "sinhalese": "si", "valencian": "ca", "panjabi": "pa", "pushto": "ps", "letzeburgesch": "lb"}

# Non-speech symbol token ids of the multilingual vocabulary (upstream tokenizer.non_speech_tokens;
# identical to the head of transformers' NON_SPEECH_TOKENS_MULTI).
NON_SPEECH_TOKENS = (
    1, 2, 7, 8, 9, 10, 14, 25, 26, 27, 28, 29, 31, 58, 59, 60, 61, 62, 63, 90, 91, 92, 93, 359, 503, 522, 542, 873, 893,
    902, 918, 922, 931, 1350, 1853, 1982, 2460, 2627, 3246, 3253, 3268, 3536, 3846, 3961, 4183, 4667, 6585, 6647, 7273,
    9061, 9383, 10428, 10929, 11938, 12033, 12331, 12562, 13793, 14157, 14635, 15265, 15618, 16553, 16604, 18362, 18956,
    20075, 21675, 22520, 26130, 26161, 26435, 28279, 29464, 31650, 32302, 32470, 36865, 42863, 47425, 49870, 50254)


def slaney_mel_filters(n_mels: int) -> np.ndarray:
    """librosa.filters.mel(sr=16000, n_fft=400, n_mels) -- Slaney scale, Slaney norm; fp32 [n_mels, 201]."""
    def hz2mel(f):
        f = np.asarray(f, np.float64)
        return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) / (np.log(6.4) / 27.0), f * 3.0 / 200.0)

    def mel2hz(m):
        m = np.asarray(m, np.float64)
        return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), m * 200.0 / 3.0)

    freqs = np.linspace(0.0, SAMPLE_RATE / 2, N_FFT // 2 + 1)
    pts = mel2hz(np.linspace(hz2mel(0.0), hz2mel(SAMPLE_RATE / 2), n_mels + 2))
    lower = (freqs[None, :] - pts[:-2, None]) / (pts[1:-1] - pts[:-2])[:, None]
    upper = (pts[2:, None] - freqs[None, :]) / (pts[2:] - pts[1:-1])[:, None]
    fb = np.maximum(0.0, np.minimum(lower, upper)) * (2.0 / (pts[2:] - pts[:-2]))[:, None]
    return fb.astype(np.float32)


class Tokens:
    """Special-token ids (upstream tokenizer.py); ids, not strings, are what the device needs."""

    def __init__(self, n_vocab: int, language: str = "ja", task: str = "transcribe"):
        self.n_vocab = n_vocab
        self.eot, self.sot = 50257, 50258
        self.num_languages = n_vocab - 51765 - 1
        base = self.sot + 1 + self.num_languages
        self.translate, self.transcribe, self.sot_lm, self.sot_prev = base, base + 1, base + 2, base + 3
        self.no_speech, self.no_timestamps, self.timestamp_begin = base + 4, base + 5, base + 6
        language = LANGUAGE_ALIASES.get(str(language).lower(), str(language).lower())
        if language not in LANGUAGES[: self.num_languages]:
            raise ValueError(f"Unsupported language: {language}")
        self.language = language
        if task not in ("transcribe", "translate"):
            raise ValueError(f"unsupported task {task!r}")
        self.language_token = self.sot + 1 + LANGUAGES.index(language)
        self.task_token = self.transcribe if task == "transcribe" else self.translate
        self.blank = 220

    def sot_sequence(self, without_timestamps: bool) -> List[int]:
        seq = [self.sot, self.language_token, self.task_token]
        return seq + [self.no_timestamps] if without_timestamps else seq

    def suppress_list(self, spec) -> List[int]:
        if isinstance(spec, str):
            spec = [int(t) for t in spec.split(",")]
        spec = list(spec or [])
        if -1 in spec:
            spec = [t for t in spec if t >= 0] + list(NON_SPEECH_TOKENS)
        spec += [self.transcribe, self.translate, self.sot, self.sot_prev, self.sot_lm, self.no_speech]
        return sorted(set(spec))


_detok_hook = None
_warned = set()


def _warn_once(key: str, msg: str) -> None:
    if key not in _warned:
        _warned.add(key)
        import logging
        logging.getLogger("whisperjav_b200").warning(msg)



def set_detokenizer(fn) -> None:
    """Install a real ``ids -> str`` detokenizer (e.g. a tiktoken/HF tokenizer when its vocabulary file
    is available).  Without one a deterministic placeholder maps every id to one CJK code point."""
    global _detok_hook
    _detok_hook = fn


def detokenize(ids: Sequence[int]) -> str:
    if _detok_hook is not None:
        return _detok_hook(list(ids))
    return "".join(chr(0x4E00 + (int(t) % 20992)) for t in ids)


def detokenize_with_specials(tok: "Tokens"):
    """``tokenizer.decode_with_timestamps`` stand-in for the word splitter: text ids through the detokenizer, EOT as its tag."""
    def decode(ids):
        out, run = [], []
        for t in ids:
            if t >= tok.eot:
                if run:
                    out.append(detokenize(run))
                    run = []
                out.append("<|endoftext|>" if t == tok.eot else f"<|{t}|>")
            else:
                run.append(t)
        if run:
            out.append(detokenize(run))
        return "".join(out)
    return decode


def compression_ratio(text: str) -> float:
    raw = text.encode("utf-8")
    return len(raw) / len(zlib.compress(raw))


@dataclass
class DecodingResult:
    tokens: List[int] = field(default_factory=list)
    text: str = ""
    avg_logprob: float = float("nan")
    no_speech_prob: float = float("nan")
    temperature: float = 0.0
    compression_ratio: float = float("nan")
    language: str = "ja"
    sum_logprob: float = float("nan")
    complete: bool = True     # False: the run stopped at its step cap before this window had ended (EOT / enough finished beams)
    steps_needed: int = 0     # sampling steps the window took (its share of a device pass)


_DECODE_KEYS = {"task", "language", "temperature", "sample_len", "best_of", "beam_size", "patience", "length_penalty",
                "prompt", "prefix", "suppress_tokens", "suppress_blank", "without_timestamps", "max_initial_timestamp", "fp16"}


class StepCapPlanner:
    """Tiered decoding for calls that span several device passes.  A pass costs as many decoder steps as its slowest window, and
    window lengths are not known in advance; so a first-tier pass is stopped after ``c1`` sampling steps, the windows that have
    ended by then are final (greedy and beam search at T = 0 are deterministic: stopping the pass does not change them) and the few
    that have not are pooled and decoded again, from scratch, in passes of their own -- stopped in turn at ``c2``, whose leftovers
    run uncapped in a last tier.  With f(c) the share of windows longer than c, a window costs about
    (c1 + f(c1) (c2 + o) + f(c2) (full + o)) / max_batch steps (o = the fixed cost of a pass): the planner picks the ladder (no cap, [c1] or [c1, c2]) that minimises it on the
    lengths seen so far (the first pass runs uncapped; a window cut off counts as ``full`` until a later tier tells its length), and
    keeps passes uncapped when that saves less than 10 %."""
    CANDIDATES = (8, 12, 16, 24, 32, 48, 64, 96, 128, 160)
    PASS_OVERHEAD = 8   # what a device pass costs besides its decoder steps (cross-K/V projection, graph capture, read-back), in steps

    def __init__(self, full: int, min_obs: int = 32):
        self.full, self.min_obs = int(full), int(min_obs)
        self.obs: List[int] = []        # steps needed per window of the first tier (a window cut off counts as `full` ...)
        self.cut = 0                    # ... and is counted here until resolve() replaces it by its length
        self.pass_max: List[int] = []

    def observe(self, results: Sequence["DecodingResult"]) -> None:
        """the results of a first-tier pass"""
        steps = [r.steps_needed if r.complete else self.full for r in results]
        if steps:
            self.obs.extend(steps)
            self.cut += sum(1 for r in results if not r.complete)
            self.pass_max.append(max(steps))

    def resolve(self, results: Sequence["DecodingResult"]) -> None:
        """windows of a later tier that have ended: their lengths replace as many ``full`` placeholders"""
        for r in results:
            if r.complete and self.cut > 0:
                self.obs.remove(self.full)
                self.obs.append(min(int(r.steps_needed), self.full))
                self.cut -= 1

    def _share_longer(self, c: int) -> float:
        return sum(1 for x in self.obs if x > c) / float(len(self.obs))

    def ladder(self) -> List[int]:
        if len(self.obs) < self.min_obs:
            return []
        best: List[int] = []
        best_cost = float(sum(self.pass_max)) / len(self.pass_max) * 0.9
        cands = [c for c in self.CANDIDATES if c < self.full]
        share = {c: self._share_longer(c) for c in cands}
        o = self.PASS_OVERHEAD
        for i, c1 in enumerate(cands):
            cost = c1 + share[c1] * (self.full + o)
            if cost < best_cost:
                best, best_cost = [c1], cost
            for c2 in cands[i + 1:]:
                cost = c1 + share[c1] * (c2 + o) + share[c2] * (self.full + o)
                if cost < best_cost * 0.97:   # a third tier has to pay for more than the model sees
                    best, best_cost = [c1, c2], cost
        return best

    def cap(self, tier: int = 0) -> Optional[int]:
        lad = self.ladder()
        return lad[tier] if tier < len(lad) else None


class TierScheduler:
    """The pass bookkeeping of tiered decoding, free of device specifics: ``decode(rows, idx, cap) -> results`` runs one device pass
    over the given encoder rows (``cap`` sampling steps at most, None = to the end), ``finish(idx, rows, sizes, results)`` consumes the
    windows that have ended.  Windows a capped pass cut off wait, with their encoder rows, in the pool of the next tier; a pool is
    decoded as soon as it holds a full batch, and ``drain()`` empties the pools in tier order."""
    MAX_CAPPED_TIERS = 2   # whatever the planner says, the third tier runs to the end: every window finishes

    def __init__(self, planner: Optional[StepCapPlanner], max_batch: int, decode, finish, stats: Optional[dict] = None):
        self.planner, self.max_batch, self.decode, self.finish = planner, int(max_batch), decode, finish
        self.pools: Dict[int, List[Tuple[int, torch.Tensor, int]]] = {}
        self.stats = stats if stats is not None else {}

    def _run(self, tier: int, idx: List[int], rows: torch.Tensor, sizes: List[int]) -> None:
        cap = self.planner.cap(tier) if (self.planner is not None and tier < self.MAX_CAPPED_TIERS) else None
        results = self.decode(rows, idx, cap)
        if self.planner is not None:
            if tier == 0:
                self.planner.observe(results)
            else:
                self.planner.resolve(results)
        late = [j for j, r in enumerate(results) if not r.complete] if cap is not None else []
        if not late:
            self.finish(idx, rows, sizes, results)
            return
        self.stats["windows_redecoded"] = self.stats.get("windows_redecoded", 0) + len(late)
        self.pools.setdefault(tier + 1, []).extend((idx[j], rows[j].clone(), sizes[j]) for j in late)
        keep = [j for j in range(len(idx)) if results[j].complete]
        if keep:
            self.finish([idx[j] for j in keep], rows[torch.tensor(keep, device=rows.device)].contiguous(), [sizes[j] for j in keep],
                        [results[j] for j in keep])

    def _flush(self, tier: int, k: int) -> None:
        pool = self.pools[tier]
        take, self.pools[tier] = pool[:k], pool[k:]
        self._run(tier, [i for i, _, _ in take], torch.stack([x for _, x, _ in take]), [z for _, _, z in take])

    def submit(self, idx: List[int], rows: torch.Tensor, sizes: List[int]) -> None:
        """one first-tier pass, then every pool that has filled up"""
        self._run(0, idx, rows, sizes)
        tier = 1
        while tier in self.pools:
            while len(self.pools[tier]) >= self.max_batch:
                self._flush(tier, self.max_batch)
            tier += 1

    def drain(self) -> None:
        tier = 1
        while tier in self.pools:
            while self.pools[tier]:
                self._flush(tier, min(len(self.pools[tier]), self.max_batch))
            tier += 1


class WhisperB200:
    """Weights resident on one B200 + reusable workspaces; ``transcribe`` mirrors upstream's signature."""
    tiered_decode = True   # StepCapPlanner for calls with more windows than max_batch (results are identical either way)
    trace_timing = False   # transcribe_batch prints per-stage wall times (adds a device synchronisation after every stage)

    def __init__(self, dims: Dims, state_dict: Dict[str, torch.Tensor], device: Union[str, int, torch.device] = "cuda",
                 max_batch: int = 64):
        if not torch.cuda.is_available():
            raise _lib.WjbError("WhisperB200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.dims = dims
        self.device = torch.device(device if device != "cuda" else f"cuda:{torch.cuda.current_device()}")
        self.max_batch = int(max_batch)
        self.is_multilingual = dims.n_vocab >= 51865
        with torch.cuda.device(self.device):
            self._blob = pack_weights(dims, state_dict, self.device)
            self._cdims = _lib.make_dims(dims)
            h = C.c_void_p()
            _lib.check(self.lib.wjb_model_create(C.byref(self._cdims), _lib.ptr(self._blob), C.byref(h)), "wjb_model_create")
            self._h = h
            self._filters = torch.from_numpy(slaney_mel_filters(dims.n_mels)).to(self.device)
        self._bufs: Dict[str, torch.Tensor] = {}
        self._pinned: Dict[str, torch.Tensor] = {}
        self.stats = {"windows": 0, "decode_steps": 0, "device_passes": 0}
        self._sample_calls = 0

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            self.lib.wjb_model_destroy(self._h)
            self._h = None
        self._bufs.clear()
        self._blob = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _buf(self, key: str, nbytes: int) -> torch.Tensor:
        t = self._bufs.get(key)
        if t is None or t.numel() < nbytes:
            t = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self._bufs[key] = t
        return t

    # ------------------------------------------------------------------ stages
    def log_mel(self, audio: torch.Tensor, n_samples: torch.Tensor, n_frames: int = N_FRAMES, layout: str = "time",
                reflect_total: int = 0, reuse: Optional[str] = None) -> torch.Tensor:
        """audio fp32 [B, S] on device, n_samples int32 [B] on device.
        layout "time": fp16 [B, n_frames + 2, n_mels] (rows 0 / n_frames+1 are the conv zero pad);
        layout "mel":  fp16 [B, n_mels, n_frames] (upstream layout).
        ``reuse``: name of a cached output buffer (the transcribe loop's; the result is then only valid until the next call
        with the same name) instead of a fresh allocation."""
        B = audio.shape[0]
        m = self.dims.n_mels
        ws = self._buf("mel_ws", self.lib.wjb_logmel_workspace_bytes(B, m))
        if layout == "time":
            if reuse:
                out = self._buf(reuse, B * (n_frames + 2) * m * 2)[: B * (n_frames + 2) * m * 2].view(torch.float16).view(B, n_frames + 2, m)
                out[:, 0].zero_()       # the kernel writes every frame row (zeros past the content); only the two conv pad
                out[:, -1].zero_()      # rows need clearing
            else:
                out = torch.zeros(B, n_frames + 2, m, dtype=torch.float16, device=self.device)
            tm, stride, row0 = 1, (n_frames + 2) * m, 1
        else:
            out = torch.empty(B, m, n_frames, dtype=torch.float16, device=self.device)
            tm, stride, row0 = 0, m * n_frames, 0
        with torch.cuda.device(self.device):
            _lib.check(self.lib.wjb_logmel_f16(_lib.ptr(audio), audio.stride(0), _lib.ptr(n_samples), B, m, _lib.ptr(self._filters),
                                              _lib.ptr(out), tm, stride, row0, n_frames, reflect_total, _lib.ptr(ws),
                                              _lib.stream_ptr()), "wjb_logmel_f16")
        return out

    def encode(self, mel_tm: torch.Tensor, tap_every: int = 0, reuse: Optional[str] = None):
        """mel_tm fp16 [B, 3002, n_mels] -> audio features fp16 [B, 1500, n_state].  ``tap_every`` (parity tests): also return
        the residual stream after every ``tap_every``-th block, fp16 [n_audio_layer // tap_every, B, 1500, n_state].
        ``reuse``: name of a cached output buffer (valid until the next call with that name)."""
        B = mel_tm.shape[0]
        d = self.dims
        if tap_every:
            taps = torch.empty(d.n_audio_layer // tap_every, B, d.n_audio_ctx, d.n_audio_state, dtype=torch.float16, device=self.device)
            _lib.check(self.lib.wjb_encoder_set_tap(self._h, _lib.ptr(taps), tap_every), "wjb_encoder_set_tap")
            try:
                return self.encode(mel_tm), taps
            finally:
                self.lib.wjb_encoder_set_tap(self._h, None, 0)
        assert mel_tm.shape[1] == 2 * d.n_audio_ctx + 2 and mel_tm.shape[2] == d.n_mels and mel_tm.is_contiguous()
        if reuse:
            nbytes = B * d.n_audio_ctx * d.n_audio_state * 2
            out = self._buf(reuse, nbytes)[:nbytes].view(torch.float16).view(B, d.n_audio_ctx, d.n_audio_state)
        else:
            out = torch.empty(B, d.n_audio_ctx, d.n_audio_state, dtype=torch.float16, device=self.device)
        nb = self.lib.wjb_encoder_workspace_bytes(self._h, B)
        ws = self._buf("enc_ws", nb)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.wjb_encoder_forward(self._h, _lib.ptr(mel_tm), B, _lib.ptr(out), _lib.ptr(ws), nb, _lib.stream_ptr()),
                       "wjb_encoder_forward")
        return out

    def decode_features(self, xa: torch.Tensor, *, language="ja", task="transcribe", without_timestamps=False,
                        suppress_tokens="-1", suppress_blank=True, max_initial_timestamp: Optional[float] = 1.0,
                        sample_len: Optional[int] = None, prompt: Optional[Sequence[int]] = None,
                        prefix: Optional[Sequence[int]] = None, temperature: float = 0.0, seed: int = 0,
                        beam_size: Optional[int] = None, patience: Optional[float] = None,
                        length_penalty: Optional[float] = None, _trace: Optional[dict] = None) -> List[DecodingResult]:
        """Decode B windows (upstream DecodingTask.run) in one device-resident loop: GreedyDecoder (argmax at T == 0,
        Categorical(logits / T) otherwise) or, with ``beam_size``, BeamSearchDecoder at T == 0."""
        d = self.dims
        B = xa.shape[0]
        tok = Tokens(d.n_vocab, language, task)
        n_ctx = d.n_text_ctx
        sample_len = sample_len or n_ctx // 2
        initial = tok.sot_sequence(without_timestamps)
        if prefix:
            max_prefix_len = n_ctx // 2 - sample_len
            initial = initial + list(prefix)[-max_prefix_len:]  # upstream slices unconditionally ([-0:] keeps everything)
        if prompt:
            initial = [tok.sot_prev] + list(prompt)[-(n_ctx // 2 - 1):] + initial
        n_initial = len(initial)
        sample_len = min(sample_len, n_ctx - n_initial + 1)
        stride = n_initial + sample_len + 1
        opts = _lib.DecodeOpts()
        opts.n_initial, opts.sot_index, opts.sample_len = n_initial, initial.index(tok.sot), sample_len
        opts.eot, opts.no_speech, opts.no_timestamps, opts.timestamp_begin = tok.eot, tok.no_speech, tok.no_timestamps, tok.timestamp_begin
        opts.suppress_blank, opts.blank_token = int(bool(suppress_blank)), tok.blank
        opts.apply_timestamp_rules = int(not without_timestamps)
        mi = -1
        if not without_timestamps and max_initial_timestamp:
            mi = round(max_initial_timestamp / (30.0 / d.n_audio_ctx))
        opts.max_initial_timestamp_index = mi
        opts.tokens_stride, opts.check_every = stride, 8
        opts.temperature, opts.seed = float(temperature), int(seed) & 0xFFFFFFFF

        # upstream _get_suppress_tokens: None / "" / [] still suppress the task and sot-family specials and no_speech
        key = ("mask", str(suppress_tokens), language, task)
        mask = self._bufs.get(key)
        if mask is None:
            mk = torch.zeros(d.n_vocab, dtype=torch.uint8)
            mk[tok.suppress_list(suppress_tokens if suppress_tokens not in (None, "") else [])] = 1
            mask = mk.to(self.device)
            self._bufs[key] = mask
        if beam_size:
            if temperature != 0.0:
                raise ValueError("beam search runs at temperature 0 (upstream drops beam_size when t > 0)")
            return self._decode_beam(xa, opts, mask, initial, tok, int(beam_size), patience, length_penalty, language)
        with torch.cuda.device(self.device):
            kv_bytes = self.lib.wjb_cross_kv_bytes(self._h, B)
            kv = self._buf("cross_kv", kv_bytes)
            _lib.check(self.lib.wjb_cross_kv(self._h, _lib.ptr(xa), B, _lib.ptr(kv), _lib.stream_ptr()), "wjb_cross_kv")
            ws_bytes = self.lib.wjb_decode_workspace_bytes(self._h, B)
            ws = self._buf("dec_ws", ws_bytes)
            out = self._buf("dec_out", B * (stride * 4 + 16)).view(torch.int32)
            tokens = out[: B * stride].view(B, stride)
            slp = out[B * stride: B * stride + B].view(torch.float32)
            nsp = out[B * stride + B: B * stride + 2 * B].view(torch.float32)
            olen = out[B * stride + 2 * B: B * stride + 3 * B]
            tokens.zero_()
            tokens[:, :n_initial] = torch.tensor(initial, dtype=torch.int32, device=self.device)
            steps = C.c_int(0)
            if _trace is not None:  # parity-test hook: per-step raw logits, the device's own picks, optional teacher forcing
                ls = self.lib.wjb_decode_logits_stride(self._h)
                t_logits = torch.zeros(n_initial - 1 + sample_len, B, ls, dtype=torch.float16, device=self.device)
                t_sampled = torch.full((B, stride), -1, dtype=torch.int32, device=self.device)
                t_forced = None
                if _trace.get("forced") is not None:
                    ft = torch.full((B, stride), tok.eot, dtype=torch.int32)
                    for b, seq in enumerate(_trace["forced"]):
                        seq = list(seq)[: stride - n_initial]
                        ft[b, n_initial: n_initial + len(seq)] = torch.tensor(seq, dtype=torch.int32)
                    t_forced = ft.to(self.device)
                _lib.check(self.lib.wjb_decode_set_trace(self._h, _lib.ptr(t_logits), t_logits.numel() * 2, _lib.ptr(t_sampled),
                                                        _lib.ptr(t_forced)), "wjb_decode_set_trace")
            try:
                _lib.check(self.lib.wjb_decode_greedy(self._h, _lib.ptr(kv), B, C.byref(opts), _lib.ptr(mask), _lib.ptr(tokens),
                                                     _lib.ptr(slp), _lib.ptr(nsp), _lib.ptr(olen), _lib.ptr(ws), ws_bytes,
                                                     C.byref(steps), _lib.stream_ptr()), "wjb_decode_greedy")
            finally:
                if _trace is not None:
                    self.lib.wjb_decode_set_trace(self._h, None, 0, None, None)
            if _trace is not None:
                _trace["logits"] = t_logits[: steps.value, :, : d.n_vocab].float().cpu()  # [steps_run][B][V]
                _trace["sampled"] = t_sampled.cpu().numpy()
                _trace["n_initial"], _trace["steps"] = n_initial, steps.value
            host = out[: B * stride + 3 * B].cpu()  # the one device->host read of the decode
        self.stats["decode_steps"] += steps.value
        self.stats["windows"] += B
        self.stats["device_passes"] += 1
        h_tokens = host[: B * stride].view(B, stride).numpy()
        h_slp = host[B * stride: B * stride + B].view(torch.float32).numpy()
        h_nsp = host[B * stride + B: B * stride + 2 * B].view(torch.float32).numpy()
        h_len = host[B * stride + 2 * B: B * stride + 3 * B].numpy()
        results = []
        for b in range(B):
            ids = [int(t) for t in h_tokens[b, n_initial: n_initial + int(h_len[b])]]
            text = detokenize([t for t in ids if t < tok.eot]).strip()
            results.append(DecodingResult(tokens=ids, text=text, avg_logprob=float(h_slp[b]) / (len(ids) + 1),
                                          no_speech_prob=float(h_nsp[b]), temperature=float(temperature),
                                          compression_ratio=compression_ratio(text) if text else 0.0, language=language,
                                          sum_logprob=float(h_slp[b]), complete=len(ids) < sample_len,
                                          steps_needed=min(len(ids) + 1, sample_len)))
        return results

    def decode_trace(self, xa: torch.Tensor, forced_tokens: Optional[Sequence[Sequence[int]]] = None, **kw):
        """Parity-test entry: greedy decode of ``xa`` that also returns the raw step logits.  Returns (results, trace) with
        trace["logits"] fp32 [steps][B][n_vocab] (row ``n_initial - 1 + i`` holds the logits the i-th sampled token was drawn
        from), trace["sampled"] the ids the device picked at every position, trace["n_initial"].  With ``forced_tokens`` the
        device is fed those ids instead of its own picks (teacher forcing)."""
        trace = {"forced": forced_tokens}
        res = self.decode_features(xa, _trace=trace, **kw)
        return res, trace

    def align_windows(self, xa: torch.Tensor, text_tokens: Sequence[Sequence[int]], num_frames: Sequence[int], *, language="ja",
                      task="transcribe", medfilt_width: int = 7, return_matrix: bool = False, mode: str = "prefill"):
        """timing.py::find_alignment for B windows in one device pass (wjb_decode_set_align + wjb_align_dtw): the decoder is
        teacher-forced over [sot sequence, <|notimestamps|>, text tokens, <|endoftext|>], the cross-attention scores of the
        alignment heads are captured, turned into the token x frame matrix and aligned by DTW.  Per window returns
        (jump_frames int array [n_text + 1], token_probs float array [n_text]); windows without text tokens get empty arrays.
        ``mode``: "prefill" = one pass with every position a GEMM row (``wjb_align_prefill``: one sweep of the decoder weights);
        "steps" = the same arithmetic through the decode step graph, one position per step (``wjb_decode_set_align``; kept as
        the cross-check of the prefill kernels, tests/test_gpu_timing.py)."""
        d = self.dims
        B = xa.shape[0]
        tok = Tokens(d.n_vocab, language, task)
        sot = [tok.sot, tok.language_token, tok.task_token]
        n_initial = len(sot) + 1
        seqs = [sot + [tok.no_timestamps] + [int(t) for t in tt] + [tok.eot] for tt in text_tokens]
        live = [b for b in range(B) if len(text_tokens[b]) > 0]
        empty = (np.zeros(0, np.int32), np.zeros(0, np.float32))
        if not live:
            return [empty] * B
        if len(live) < B:
            sub = self.align_windows(xa[torch.tensor(live, device=self.device)].contiguous(), [text_tokens[b] for b in live],
                                     [num_frames[b] for b in live], language=language, task=task, medfilt_width=medfilt_width,
                                     return_matrix=return_matrix, mode=mode)
            out = [empty + ((None,) if return_matrix else ())] * B
            for b, r in zip(live, sub):
                out[b] = r
            return out
        max_len = max(len(q) for q in seqs)
        if max_len > d.n_text_ctx:
            raise ValueError("alignment sequence longer than n_text_ctx")
        if mode == "prefill" and max_len <= 256:
            max_len = (max_len + 7) // 8 * 8  # padded row count per window
        else:
            mode = "steps"
        sample_len = max_len - n_initial + 1
        stride = n_initial + sample_len + 1
        opts = _lib.DecodeOpts()
        opts.n_initial, opts.sot_index, opts.sample_len = n_initial, 0, sample_len
        opts.eot, opts.no_speech, opts.no_timestamps, opts.timestamp_begin = tok.eot, tok.no_speech, tok.no_timestamps, tok.timestamp_begin
        opts.suppress_blank, opts.blank_token, opts.apply_timestamp_rules, opts.max_initial_timestamp_index = 0, tok.blank, 0, -1
        opts.tokens_stride, opts.check_every, opts.temperature, opts.seed = stride, 8, 0.0, 0
        dev = self.device
        with torch.cuda.device(dev):
            kv = self._buf("cross_kv", self.lib.wjb_cross_kv_bytes(self._h, B))
            _lib.check(self.lib.wjb_cross_kv(self._h, _lib.ptr(xa), B, _lib.ptr(kv), _lib.stream_ptr()), "wjb_cross_kv")
            ws_bytes = self.lib.wjb_decode_workspace_bytes(self._h, B)
            ws = self._buf("dec_ws", ws_bytes)
            tokens = torch.zeros(B, stride, dtype=torch.int32)
            forced = torch.full((B, stride), tok.eot, dtype=torch.int32)
            for b, q in enumerate(seqs):
                tokens[b, :n_initial] = torch.tensor(q[:n_initial], dtype=torch.int32)
                forced[b, : len(q)] = torch.tensor(q, dtype=torch.int32)
            tokens, forced = tokens.to(dev), forced.to(dev)
            n_tok = torch.tensor([len(q) for q in seqs], dtype=torch.int32, device=dev)
            row_begin = torch.full((B,), len(sot), dtype=torch.int32, device=dev)
            n_rows = torch.tensor([len(q) - len(sot) - 1 for q in seqs], dtype=torch.int32, device=dev)
            nf2 = torch.tensor([int(n) // 2 for n in num_frames], dtype=torch.int32, device=dev)
            qk = self._buf("align_qk", self.lib.wjb_align_qk_bytes(self._h, B, max_len))
            prob = torch.zeros(B, stride, dtype=torch.float32, device=dev)
            slp = torch.zeros(3, B, dtype=torch.float32, device=dev)
            olen = torch.zeros(B, dtype=torch.int32, device=dev)
            steps = C.c_int(0)
            if mode == "prefill":
                pws_bytes = self.lib.wjb_align_prefill_workspace_bytes(self._h, B, max_len)
                pws = self._buf("align_pf_ws", pws_bytes)
                _lib.check(self.lib.wjb_align_prefill(self._h, _lib.ptr(kv), _lib.ptr(forced), stride, _lib.ptr(n_tok), B, max_len, tok.eot, _lib.ptr(qk),
                                                     _lib.ptr(prob), _lib.ptr(pws), pws_bytes, _lib.stream_ptr()), "wjb_align_prefill")
            else:
                _lib.check(self.lib.wjb_decode_set_trace(self._h, None, 0, None, _lib.ptr(forced)), "wjb_decode_set_trace")
                _lib.check(self.lib.wjb_decode_set_align(self._h, _lib.ptr(qk), max_len, _lib.ptr(n_tok), _lib.ptr(prob)), "wjb_decode_set_align")
                try:
                    _lib.check(self.lib.wjb_decode_greedy(self._h, _lib.ptr(kv), B, C.byref(opts), None, _lib.ptr(tokens), _lib.ptr(slp[0]),
                                                         _lib.ptr(slp[1]), _lib.ptr(olen), _lib.ptr(ws), ws_bytes, C.byref(steps),
                                                         _lib.stream_ptr()), "wjb_decode_greedy (alignment pass)")
                finally:
                    self.lib.wjb_decode_set_align(self._h, None, 0, None, None)
                    self.lib.wjb_decode_set_trace(self._h, None, 0, None, None)
            matrix = torch.zeros(B, max_len, d.n_audio_ctx, dtype=torch.float32, device=dev)
            jump = torch.zeros(B, max_len, dtype=torch.int32, device=dev)
            aws_bytes = self.lib.wjb_align_workspace_bytes(self._h, B, max_len)
            aws = self._buf("align_ws", aws_bytes)
            _lib.check(self.lib.wjb_align_dtw(self._h, _lib.ptr(qk), B, max_len, _lib.ptr(n_tok), _lib.ptr(row_begin), _lib.ptr(n_rows),
                                             _lib.ptr(nf2), int(medfilt_width), _lib.ptr(matrix), _lib.ptr(jump), _lib.ptr(aws), aws_bytes,
                                             _lib.stream_ptr()), "wjb_align_dtw")
            h_jump, h_prob = jump.cpu().numpy(), prob.cpu().numpy()
        self.stats["decode_steps"] += steps.value
        self.stats["device_passes"] += 1
        out = []
        for b, q in enumerate(seqs):
            n_text = len(q) - n_initial - 1
            r = (h_jump[b, : n_text + 1].copy(), h_prob[b, n_initial: n_initial + n_text].copy())
            if return_matrix:
                r = r + (matrix[b, len(sot): len(q) - 1, : int(num_frames[b]) // 2].cpu(),)
            out.append(r)
        return out

    def _decode_beam(self, xa, opts, mask, initial, tok, beam: int, patience, length_penalty, language) -> List[DecodingResult]:
        """upstream decoding.py::DecodingTask.run with BeamSearchDecoder: the device runs update() for every step
        (``wjb_decode_beam``: rows = windows x beams, cache ancestry tables instead of cache permutation); finalize() and the
        MaximumLikelihoodRanker are the host logic below, on one device->host read."""
        d = self.dims
        n_audio = xa.shape[0]
        rows = n_audio * beam
        max_cand = round(beam * (patience or 1.0))
        if max_cand <= 0:
            raise ValueError(f"Invalid beam size ({beam}) or patience ({patience})")
        n_initial, stride, n_ctx = opts.n_initial, opts.tokens_stride, d.n_text_ctx
        dev = self.device
        with torch.cuda.device(dev):
            kv_bytes = self.lib.wjb_cross_kv_bytes(self._h, n_audio)
            kv = self._buf("cross_kv", kv_bytes)
            _lib.check(self.lib.wjb_cross_kv(self._h, _lib.ptr(xa), n_audio, _lib.ptr(kv), _lib.stream_ptr()), "wjb_cross_kv")
            ws_bytes = self.lib.wjb_decode_workspace_bytes(self._h, rows)
            ws = self._buf("dec_ws", ws_bytes)
            tokens = torch.zeros(2, rows, stride, dtype=torch.int32, device=dev)
            tokens[:, :, :n_initial] = torch.tensor(initial, dtype=torch.int32, device=dev)
            anc = torch.arange(rows, dtype=torch.int16, device=dev).view(1, rows, 1).expand(2, rows, n_ctx).contiguous()
            slp = torch.zeros(2, rows, dtype=torch.float32, device=dev)
            fin_tokens = torch.zeros(n_audio, max_cand, stride, dtype=torch.int32, device=dev)
            fin_score = torch.zeros(n_audio, max_cand, dtype=torch.float32, device=dev)
            fin_len = torch.zeros(n_audio, max_cand, dtype=torch.int32, device=dev)
            fin_count = torch.zeros(n_audio, dtype=torch.int32, device=dev)
            audio_done = torch.zeros(n_audio, dtype=torch.uint8, device=dev)
            nsp = torch.zeros(n_audio, dtype=torch.float32, device=dev)
            bufs = _lib.BeamBufs(n_audio, beam, max_cand, _lib.ptr(tokens), _lib.ptr(anc), _lib.ptr(slp), _lib.ptr(fin_tokens),
                                 _lib.ptr(fin_score), _lib.ptr(fin_len), _lib.ptr(fin_count), _lib.ptr(audio_done))
            steps = C.c_int(0)
            _lib.check(self.lib.wjb_decode_beam(self._h, _lib.ptr(kv), C.byref(bufs), C.byref(opts), _lib.ptr(mask), _lib.ptr(nsp),
                                                _lib.ptr(ws), ws_bytes, C.byref(steps), _lib.stream_ptr()), "wjb_decode_beam")
            par = steps.value & 1  # live rows after the last executed step
            h_live = tokens[par].cpu().numpy()
            h_slp = slp[par].cpu().numpy()
            h_fin_tok, h_fin_score = fin_tokens.cpu().numpy(), fin_score.cpu().numpy()
            h_fin_len, h_fin_count = fin_len.cpu().numpy(), fin_count.cpu().numpy()
            h_nsp = nsp.cpu().numpy()
            h_done = audio_done.cpu().numpy()
        self.stats["decode_steps"] += steps.value
        self.stats["windows"] += n_audio
        self.stats["device_passes"] += 1
        live_len = min(steps.value + 1, stride)  # tokens per live row: the prompt plus one per executed sampling step
        results = []
        for a in range(n_audio):
            finished = [(h_fin_tok[a, k, : int(h_fin_len[a, k])], float(h_fin_score[a, k])) for k in range(int(h_fin_count[a]))]
            live = [(h_live[a * beam + j, :live_len], float(h_slp[a * beam + j])) for j in range(beam)]
            ids, score = beam_finalize_and_rank(finished, live, beam, n_initial, tok.eot, length_penalty)
            text = detokenize([t for t in ids if t < tok.eot]).strip()
            need = max([int(h_fin_len[a, k]) for k in range(int(h_fin_count[a]))] + [n_initial]) - n_initial + 1
            results.append(DecodingResult(tokens=ids, text=text, avg_logprob=score / (len(ids) + 1), no_speech_prob=float(h_nsp[a]),
                                          temperature=0.0, compression_ratio=compression_ratio(text) if text else 0.0, language=language,
                                          sum_logprob=score, complete=bool(h_done[a]), steps_needed=min(need, opts.sample_len)))
        return results

    # ------------------------------------------------------------------ upstream-shaped API
    def transcribe(self, audio: Union[np.ndarray, torch.Tensor], **params) -> dict:
        """``whisper_model.transcribe(audio, **params)`` (whisper_pro_asr.py:433)."""
        return self.transcribe_batch([audio], **params)[0]

    def transcribe_batch(self, audios: Sequence[Union[np.ndarray, torch.Tensor]], *, verbose=None,
                         temperature: Union[float, Sequence[float]] = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0),
                         compression_ratio_threshold: Optional[float] = 2.4, logprob_threshold: Optional[float] = -1.0,
                         no_speech_threshold: Optional[float] = 0.6, condition_on_previous_text: bool = True,
                         initial_prompt: Optional[Sequence[int]] = None, word_timestamps: bool = False,
                         carry_initial_prompt: bool = False, prepend_punctuations: str = "\"'“¿([{-",
                         append_punctuations: str = "\"'.。,，!！?？:：”)]}、", clip_timestamps: Union[str, List[float]] = "0",
                         hallucination_silence_threshold: Optional[float] = None, pinned_audio: Optional[torch.Tensor] = None,
                         _record_windows: bool = False, **decode_options) -> List[dict]:
        """Independent clips (each: fp32 mono 16 kHz) -> one upstream-shaped result dict per clip.  Every keyword of upstream
        ``whisper.transcribe()`` is accepted by name; names upstream would reject raise TypeError as ``DecodingOptions(**kw)`` does."""
        unknown = set(decode_options) - _DECODE_KEYS
        if unknown:  # upstream: DecodingOptions(**kwargs) raises TypeError
            raise TypeError(f"transcribe() got unexpected keyword arguments {sorted(unknown)}")
        if clip_timestamps not in ("0", [0], [0.0], 0, None, []):
            _warn_once("clip_timestamps", "clip_timestamps other than the default \"0\" is not implemented: whole clips are transcribed")
        if hallucination_silence_threshold is not None and not word_timestamps:
            pass  # upstream only consults it when word_timestamps=True
        elif hallucination_silence_threshold is not None:
            _warn_once("hst", "hallucination_silence_threshold is accepted but the silence-skipping heuristic is not implemented")
        if decode_options.get("language") is None:
            _warn_once("lang", "language=None: upstream would run language detection; this backend decodes as 'ja' (the reference's "
                               "configured language, whisper_pro_asr.py:36)")
        language = decode_options.pop("language", None) or "ja"
        task = decode_options.pop("task", "transcribe")
        decode_options.pop("fp16", None)
        best_of = decode_options.pop("best_of", None) or 1
        # upstream decode_with_fallback: beam_size / patience apply at t == 0, best_of at t > 0
        beam_opts = {"beam_size": decode_options.pop("beam_size", None), "patience": decode_options.pop("patience", None),
                     "length_penalty": decode_options.pop("length_penalty", None)}
        if beam_opts["beam_size"]:
            decode_options["_beam"] = beam_opts
        temps = [temperature] if isinstance(temperature, (int, float)) else list(temperature)
        d = self.dims
        tok = Tokens(d.n_vocab, language, task)
        n = len(audios)
        arrs = [a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a) for a in audios]
        arrs = [a.astype(np.float32, copy=False).reshape(-1) for a in arrs]
        content = [len(a) // HOP_LENGTH for a in arrs]
        state = [{"seek": 0, "all_tokens": list(initial_prompt or []), "reset": 0, "segments": [], "last_speech": 0.0, "windows": []}
                 for _ in range(n)]
        init_len = len(initial_prompt or [])
        for st in state:
            st["reset"] = 0

        import time
        _trace = bool(self.trace_timing)   # diagnostic: per-stage wall times of this call on stderr (synchronises after every stage)
        _t = {"t0": time.perf_counter()}

        def _mark(name):
            if _trace:
                torch.cuda.synchronize()
                _t[name] = time.perf_counter()

        # clip-level log-mel (content frames only), computed once per clip batch
        mels = self._clip_mels(arrs, content, pinned_audio)
        _mark("mel")

        full_len = int(decode_options.get("sample_len") or d.n_text_ctx // 2)

        def finish(idx: List[int], xa_rows, sizes_rows: List[int], results: List[DecodingResult]) -> None:
            """the host half of the seek-loop iteration for the windows of one device pass"""
            if _record_windows:  # parity tests: what every decoded window returned, before the host logic touches it
                for j, i in enumerate(idx):
                    r = results[j]
                    state[i]["windows"].append({"seek": state[i]["seek"], "size": sizes_rows[j], "tokens": list(r.tokens), "avg_logprob": r.avg_logprob,
                                                "no_speech_prob": r.no_speech_prob, "temperature": r.temperature})
            pend = [self._slice(state[i], results[j], tok, sizes_rows[j], no_speech_threshold, logprob_threshold) for j, i in enumerate(idx)]
            if word_timestamps:
                self._word_timestamps(xa_rows, pend, sizes_rows, tok, language, task, prepend_punctuations, append_punctuations,
                                      [state[i] for i in idx])
            for j, i in enumerate(idx):
                self._commit(state[i], pend[j], results[j], condition_on_previous_text)

        def decode(xa_rows, idx: List[int], step_cap: Optional[int]):
            prompts = [state[i]["all_tokens"][state[i]["reset"]:] for i in idx]
            return self._decode_with_fallback(xa_rows, prompts, temps, best_of, language, task, decode_options,
                                              compression_ratio_threshold, logprob_threshold, no_speech_threshold, step_cap=step_cap)

        while True:
            active = [i for i in range(n) if state[i]["seek"] < content[i]]
            if not active:
                break
            planner = None
            if self.tiered_decode and len(active) > self.max_batch and not decode_options.get("prefix") and temps[0] == 0:
                planner = StepCapPlanner(full_len, min_obs=min(32, self.max_batch))
            sched = TierScheduler(planner, self.max_batch, decode, finish, self.stats)
            for c0 in range(0, len(active), self.max_batch):
                chunk = active[c0: c0 + self.max_batch]
                sizes = [min(N_FRAMES, content[i] - state[i]["seek"]) for i in chunk]
                win = self._gather_windows(mels, chunk, [state[i]["seek"] for i in chunk], sizes)
                _mark("gather")
                xa = self.encode(win, reuse="xa_loop")
                _mark("encode")
                sched.submit(chunk, xa, sizes)
                _mark("decode+advance")
            sched.drain()
        if _trace:
            import sys
            keys = list(_t)
            print("wjb timing (ms):", {k: round((_t[k] - _t[keys[max(0, keys.index(k) - 1)]]) * 1e3, 1) for k in keys[1:]},
                  "passes", self.stats["device_passes"], file=sys.stderr)
        outs = []
        for i in range(n):
            toks = state[i]["all_tokens"][init_len:]
            outs.append({"text": detokenize([t for t in toks if t < tok.eot]), "segments": state[i]["segments"], "language": language})
            if _record_windows:
                outs[-1]["windows"] = state[i]["windows"]
        return outs

    # -- helpers -------------------------------------------------------------------------------
    def _clip_mels(self, arrs: List[np.ndarray], content: List[int], pinned: Optional[torch.Tensor] = None):
        """Upload clips (pinned -> device) and compute their content-frame log-mel, time-major.
        ``pinned``: optional caller-owned pinned fp32 [n, S] tensor already holding the clips (skips staging)."""
        n = len(arrs)
        if n == 0:
            return None
        max_s = max(max(len(a) for a in arrs), HOP_LENGTH)
        max_f = max(max(content), 1)
        if pinned is not None and pinned.shape[0] == n and pinned.shape[1] >= max_s and pinned.is_pinned():
            host = pinned
        else:
            host = self._pinned.get("audio")
            if host is None or host.shape[0] < n or host.shape[1] < max_s:
                host = torch.zeros(n, max_s, dtype=torch.float32).pin_memory()
                self._pinned["audio"] = host
            host = host[:n, :max_s]
            for i, a in enumerate(arrs):
                host[i, : len(a)] = torch.from_numpy(a)
                host[i, len(a):] = 0
        ns = torch.tensor([len(a) for a in arrs], dtype=torch.int32)
        dev_audio = host.to(self.device, non_blocking=True)
        dev_ns = ns.to(self.device)
        # frames per clip rounded so short clips still give one full window directly
        nf = max(max_f, N_FRAMES)
        return self.log_mel(dev_audio, dev_ns, n_frames=nf, layout="time", reuse="mel_loop"), nf

    def _gather_windows(self, mels, chunk: List[int], seeks: List[int], sizes: List[int]) -> torch.Tensor:
        mel, nf = mels
        m = self.dims.n_mels
        if nf == N_FRAMES and all(s == 0 for s in seeks):
            if chunk == list(range(mel.shape[0])):
                return mel  # the whole clip batch in order: no copy
            idx = torch.tensor(chunk, device=self.device)
            return mel.index_select(0, idx).contiguous()
        win = torch.zeros(len(chunk), N_FRAMES + 2, m, dtype=torch.float16, device=self.device)
        for j, (i, s, z) in enumerate(zip(chunk, seeks, sizes)):
            win[j, 1: 1 + z] = mel[i, 1 + s: 1 + s + z]
        return win

    def _decode_with_fallback(self, xa, prompts, temps, best_of, language, task, decode_options, cr_thr, lp_thr, ns_thr,
                              step_cap: Optional[int] = None):
        """upstream transcribe.py::decode_with_fallback, batched: every window is decoded at temps[0]; windows whose result
        looks degenerate (compression ratio too high, or average log-probability too low unless the window is judged
        silent) are re-decoded as a smaller batch at the next temperature; at T > 0 ``best_of`` samples are drawn per
        window and ranked by sum_logprob / length (MaximumLikelihoodRanker with length_penalty=None)."""
        n = len(prompts)
        final: List[Optional[DecodingResult]] = [None] * n
        todo = list(range(n))
        for ti, t in enumerate(temps):
            if not todo:
                break
            sub_xa = xa if len(todo) == n else xa[torch.tensor(todo, device=self.device)].contiguous()
            sub_prompts = [prompts[i] for i in todo]
            group = best_of if (t > 0 and best_of > 1) else 1
            cands = []
            opts_t = decode_options
            if step_cap is not None and ti == 0 and t == 0:
                opts_t = dict(decode_options, sample_len=min(int(step_cap), int(decode_options.get("sample_len") or step_cap)))
            for g in range(group):
                self._sample_calls += 1
                cands.append(self._decode_with_prompts(sub_xa, sub_prompts, t, language, task, opts_t,
                                                       seed=0x5EED0000 + 7919 * self._sample_calls))
            still = []
            for k, i in enumerate(todo):
                opts_k = [c[k] for c in cands]
                best = max(opts_k, key=lambda r: r.sum_logprob / max(len(r.tokens), 1)) if group > 1 else opts_k[0]
                final[i] = best
                if opts_t is not decode_options and not best.complete:
                    continue   # cut off by the step cap: not a result yet
                needs = False
                if cr_thr is not None and best.compression_ratio > cr_thr:
                    needs = True
                if lp_thr is not None and best.avg_logprob < lp_thr:
                    needs = True
                if ns_thr is not None and best.no_speech_prob > ns_thr and lp_thr is not None and best.avg_logprob < lp_thr:
                    needs = False
                if needs and ti + 1 < len(temps):
                    still.append(i)
            todo = still
        return final

    def _decode_with_prompts(self, xa, prompts, temperature, language, task, decode_options, seed: int = 0):
        # windows with identical prompts share one device pass (the common case: no prompt at all)
        groups: Dict[tuple, List[int]] = {}
        for j, p in enumerate(prompts):
            groups.setdefault(tuple(p), []).append(j)
        results: List[Optional[DecodingResult]] = [None] * len(prompts)
        for p, idxs in groups.items():
            sub = xa if len(idxs) == len(prompts) else xa[torch.tensor(idxs, device=self.device)].contiguous()
            beam_kw = dict(decode_options.get("_beam") or {}) if temperature == 0 else {}
            res = self.decode_features(sub, language=language, task=task, prompt=list(p) or None, temperature=temperature, seed=seed, **beam_kw,
                                       **{k: v for k, v in decode_options.items() if k in
                                          ("without_timestamps", "suppress_tokens", "suppress_blank", "max_initial_timestamp",
                                           "sample_len", "prefix")})
            for j, r in zip(idxs, res):
                results[j] = r
        return results

    @staticmethod
    def _slice(st: dict, result: DecodingResult, tok: Tokens, segment_size: int, no_speech_threshold, logprob_threshold) -> dict:
        """One iteration of upstream transcribe()'s seek loop for one clip, up to the word-timestamp hook: the no-speech skip,
        timestamp-token slicing into segments and the provisional seek.  Returns {"skip", "current", "seek", "single_ending"}."""
        seek = st["seek"]
        tokens = result.tokens
        if no_speech_threshold is not None:
            should_skip = result.no_speech_prob > no_speech_threshold
            if logprob_threshold is not None and result.avg_logprob > logprob_threshold:
                should_skip = False
            if should_skip:
                return {"skip": True, "current": [], "seek": seek + segment_size, "single_ending": False}
        time_offset = float(seek * HOP_LENGTH / SAMPLE_RATE)
        precision = 0.02
        fields = {"temperature": result.temperature, "avg_logprob": result.avg_logprob,
                  "compression_ratio": result.compression_ratio, "no_speech_prob": result.no_speech_prob}

        def seg(start, end, toks):
            return {"seek": seek, "start": start, "end": end, "text": detokenize([t for t in toks if t < tok.eot]),
                    "tokens": list(toks), **fields}

        is_ts = [t >= tok.timestamp_begin for t in tokens]
        single_ending = is_ts[-2:] == [False, True]
        cuts = [k + 1 for k in range(len(tokens) - 1) if is_ts[k] and is_ts[k + 1]]
        current = []
        if cuts:
            if single_ending:
                cuts.append(len(tokens))
            last = 0
            for cut in cuts:
                piece = tokens[last:cut]
                current.append(seg(time_offset + (piece[0] - tok.timestamp_begin) * precision,
                                   time_offset + (piece[-1] - tok.timestamp_begin) * precision, piece))
                last = cut
            if single_ending:
                new_seek = seek + segment_size
            else:
                new_seek = seek + (tokens[last - 1] - tok.timestamp_begin) * 2
        else:
            duration = segment_size * HOP_LENGTH / SAMPLE_RATE
            stamps = [t for t in tokens if t >= tok.timestamp_begin]
            if stamps and stamps[-1] != tok.timestamp_begin:
                duration = (stamps[-1] - tok.timestamp_begin) * precision
            current.append(seg(time_offset, time_offset + duration, tokens))
            new_seek = seek + segment_size
        return {"skip": False, "current": current, "seek": new_seek, "single_ending": single_ending}

    def _word_timestamps(self, xa, pend: List[dict], sizes: List[int], tok: Tokens, language, task, prepend_punctuations,
                         append_punctuations, states: List[dict]) -> None:
        """upstream transcribe()'s ``if word_timestamps:`` block for every window of a device pass: one batched alignment pass
        (``align_windows``), then per window ``add_word_timestamps`` and the seek / last_speech_timestamp updates."""
        from . import timing as TM
        text = [[t for s in p["current"] for t in s["tokens"] if t < tok.eot] if not p["skip"] else [] for p in pend]
        aligned = self.align_windows(xa, text, sizes, language=language, task=task)
        for p, st, tt, (jump, probs) in zip(pend, states, text, aligned):
            if p["skip"]:
                continue
            alignment = TM.words_from_alignment(tt, jump, probs, detokenize_with_specials(tok), tok.language, tok.eot)
            st["last_speech"] = TM.add_word_timestamps(p["current"], alignment, tok.eot, prepend_punctuations, append_punctuations,
                                                       st["last_speech"])
            time_offset = float(st["seek"] * HOP_LENGTH / SAMPLE_RATE)
            if not p["single_ending"]:
                lwe = TM.last_word_end(p["current"])
                if lwe is not None and lwe > time_offset:
                    p["seek"] = round(lwe * 100)  # FRAMES_PER_SECOND
            lwe = TM.last_word_end(p["current"])
            if lwe is not None:
                st["last_speech"] = lwe

    @staticmethod
    def _advance(st: dict, result: DecodingResult, tok: Tokens, segment_size: int, no_speech_threshold, logprob_threshold,
                 condition_on_previous_text: bool) -> None:
        """One whole seek-loop iteration without word timestamps (= _slice + _commit)."""
        WhisperB200._commit(st, WhisperB200._slice(st, result, tok, segment_size, no_speech_threshold, logprob_threshold), result,
                            condition_on_previous_text)

    @staticmethod
    def _commit(st: dict, p: dict, result: DecodingResult, condition_on_previous_text: bool) -> None:
        """The tail of the seek-loop iteration: clear instantaneous / empty segments, append, advance the seek, reset the prompt."""
        st["seek"] = p["seek"]
        if p["skip"]:
            return
        current = p["current"]
        for s in current:
            if s["start"] == s["end"] or s["text"].strip() == "":
                s["text"], s["tokens"] = "", []
                if "words" in s:
                    s["words"] = []
        base = len(st["segments"])
        st["segments"].extend({"id": base + k, **s} for k, s in enumerate(current))
        st["all_tokens"].extend(t for s in current for t in s["tokens"])
        if not condition_on_previous_text or result.temperature > 0.5:
            st["reset"] = len(st["all_tokens"])


def load_model(name: str = "large-v3", device: Union[str, torch.device] = "cuda", state_dict: Optional[dict] = None,
               seed: Optional[int] = None, max_batch: int = 64) -> WhisperB200:
    """``whisper.load_model(name, device)`` stand-in (whisper_pro_asr.py:182).  With no checkpoint
    reachable (no network on either box) ``state_dict=None`` builds the seeded synthetic weights of
    that architecture; pass a real openai- or HF-named ``state_dict`` to run real weights."""
    if name not in DIMS:
        raise RuntimeError(f"Model {name} not found; available models = {sorted(DIMS)}")
    dims = DIMS[name]
    if state_dict is None:
        kw = synth_preset(name)
        if seed is not None:
            kw["seed"] = seed
        state_dict = synth_weights(dims, **kw)
    return WhisperB200(dims, state_dict, device=device, max_batch=max_batch)
