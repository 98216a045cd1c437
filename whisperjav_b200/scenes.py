"""B5 -- SceneDetector backend "b200-auditok": the reference's two-pass silence-based scene split
(whisperjav/modules/scene_detection_backends/auditok_backend.py:229-567; protocol base.py:185-250) with the per-block energy
gate on the GPU (csrc/scene.cu, ``wjb_scene_energy``) and the tokenizer state machine on the host.

Pass 1 finds chapters separated by long silences over the whole stream, pass 2 re-splits every chapter longer than
``max_duration`` with a higher threshold / shorter silence; chapters pass 2 cannot split are cut by the clock.  The
reference runs ``auditok.split`` (pure-Python per-block loop over float64 numpy energies) once over the film and once per
oversized chapter; here each pass is ONE kernel launch over all its regions (exact integer sums of squares per 50 ms block,
so the valid / silent flags are the CPU path's flags bit for bit), and only the O(blocks) state machine stays on the host.

``detect()`` works on arrays in memory (the stream pipeline hands them on without the reference's WAV round trip,
scene_detection_backends/utils.py:106-140); ``detect_scenes()`` keeps the protocol: it writes the per-scene PCM16 WAVs and
returns the reference's ``SceneDetectionResult`` when WhisperJAV is importable (a structural twin otherwise).

Not implemented: ``assist_processing`` (scipy band-pass + pydub DRC before pass 2, off by default) -- accepted, warned, ignored.
"""
from __future__ import annotations

import logging
import math
import time
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

logger = logging.getLogger("whisperjav")

ANALYSIS_WINDOW_S = 0.05   # auditok.core.DEFAULT_ANALYSIS_WINDOW (the reference never overrides it)
_EPS = 1e-10


@dataclass
class SceneConfig:
    """Field for field ``AuditokSceneConfig`` (auditok_backend.py:36-91)."""
    max_duration: float = 29.0
    min_duration: float = 0.2
    pass1_min_duration: float = 0.3
    pass1_max_duration: float = 2700.0
    pass1_max_silence: float = 1.8
    pass1_energy_threshold: int = 32
    pass2_min_duration: float = 0.3
    pass2_max_duration: Optional[float] = None
    pass2_max_silence: float = 0.94
    pass2_energy_threshold: int = 38
    assist_processing: bool = False
    brute_force_fallback: bool = True
    brute_force_chunk_s: Optional[float] = None
    pad_edges_s: float = 0.0
    verbose_summary: bool = True
    force_mono: bool = True

    def __post_init__(self):
        if self.pass2_max_duration is None:
            self.pass2_max_duration = max(self.max_duration - 1.0, self.min_duration)
        if self.brute_force_chunk_s is None:
            self.brute_force_chunk_s = self.max_duration


def config_from_kwargs(kw: dict) -> SceneConfig:
    """Legacy DynamicSceneDetector-style names incl. the ``_s`` aliases (auditok_backend.py:144-214); unknown keys are ignored."""
    def pick(*names, default=None):
        for n in names:
            if n in kw and kw[n] is not None:
                return kw[n]
        return default

    legacy_sil = float(kw.get("max_silence", 1.8))
    legacy_thr = int(kw.get("energy_threshold", 32))
    p2max = pick("pass2_max_duration_s", "pass2_max_duration")
    bf = kw.get("brute_force_chunk_s")
    return SceneConfig(
        max_duration=float(pick("max_duration_s", "max_duration", default=29.0)),
        min_duration=float(pick("min_duration_s", "min_duration", default=0.2)),
        pass1_min_duration=float(pick("pass1_min_duration_s", "pass1_min_duration", default=0.3)),
        pass1_max_duration=float(pick("pass1_max_duration_s", "pass1_max_duration", default=2700.0)),
        pass1_max_silence=float(pick("pass1_max_silence_s", "pass1_max_silence", default=legacy_sil)),
        pass1_energy_threshold=int(kw.get("pass1_energy_threshold", legacy_thr)),
        pass2_min_duration=float(pick("pass2_min_duration_s", "pass2_min_duration", default=0.3)),
        pass2_max_duration=float(p2max) if p2max is not None else None,
        pass2_max_silence=float(pick("pass2_max_silence_s", "pass2_max_silence", default=0.94)),
        pass2_energy_threshold=int(kw.get("pass2_energy_threshold", 38)),
        assist_processing=bool(kw.get("assist_processing", False)),
        brute_force_fallback=bool(kw.get("brute_force_fallback", True)),
        brute_force_chunk_s=float(bf) if bf is not None else None,
        pad_edges_s=float(kw.get("pad_edges_s", 0.0)),
        verbose_summary=bool(kw.get("verbose_summary", True)),
        force_mono=bool(kw.get("force_mono", True)),
    )


# ------------------------------------------------------------------------------------------------------
# host logic: energy flags -> tokens (auditok.core.StreamTokenizer with init_min = init_max_silence = 0)
# ------------------------------------------------------------------------------------------------------

def windows_for(duration: float, analysis_window: float, up: bool) -> int:
    """auditok.core._duration_to_nb_windows: ceil for min_dur, floor(+1e-10) for max_dur / max_silence."""
    if duration == 0:
        return 0
    x = duration / analysis_window
    return int(math.ceil(x)) if up else int(math.floor(x + _EPS))


def energy_flags(sumsq: np.ndarray, counts: np.ndarray, threshold: float) -> np.ndarray:
    """upstream's decision on exact integer inputs: 20 log10(max(sqrt(mean(x^2)), 1e-10)) >= threshold, in float64."""
    mean = sumsq.astype(np.float64) / np.maximum(counts, 1).astype(np.float64)
    db = 20.0 * np.log10(np.maximum(np.sqrt(mean), _EPS))
    return db >= float(threshold)


def tokenize_flags(valid: Sequence[bool], min_len: int, max_len: int, max_sil: int, drop_trailing_silence: bool = True,
                   strict_min: bool = False) -> List[Tuple[int, int]]:
    """-> [(first block, blocks)].  Three live states (with init_min = 0 a valid block opens a token at once); only counters are
    kept: ``n`` blocks in the open token, ``sil`` trailing silent blocks, ``start``, and whether the previous token was cut by
    ``max_len`` (a follow-up shorter than ``min_len`` is still delivered then)."""
    if max_len <= 0 or min_len <= 0 or min_len > max_len:
        raise ValueError("need 0 < min_len <= max_len")
    if max_sil >= max_len:
        raise ValueError("max_silence must be shorter than max_dur")
    IDLE, OPEN, TRAIL = 0, 1, 2
    out: List[Tuple[int, int]] = []
    state, n, sil, start, contiguous = IDLE, 0, 0, 0, False

    def close(f: int, cut: bool):
        nonlocal n, sil, start, contiguous
        if not cut and drop_trailing_silence and sil > 0:
            n = max(0, n - sil)
        if n >= min_len or (n > 0 and not strict_min and contiguous):
            out.append((start, n))
            if cut:
                start = f + 1
            contiguous = cut
        else:
            contiguous = False
        n = 0

    for f, v in enumerate(valid):
        if state == IDLE:
            if v:
                start, n, sil, state = f, 1, 0, OPEN
                if n >= max_len:
                    close(f, True)
        elif state == OPEN:
            if v:
                n += 1
                if n >= max_len:
                    close(f, True)
            elif max_sil <= 0:
                state = IDLE
                close(f, False)
            else:
                sil, state = 1, TRAIL
                n += 1
                if n == max_len:
                    close(f, True)
        else:  # TRAIL
            if v:
                n += 1
                sil, state = 0, OPEN
                if n >= max_len:
                    close(f, True)
            elif sil >= max_sil:
                state = IDLE
                if sil < n:
                    close(f, False)
                else:
                    n, sil = 0, 0
            else:
                n += 1
                sil += 1
                if n >= max_len:
                    close(f, True)
    if state != IDLE and n > 0 and n > sil:
        close(len(valid) - 1, False)
    return out


def split_region(sumsq: np.ndarray, n_samples: int, window: int, sr: int, min_dur: float, max_dur: float, max_silence: float,
                 energy_threshold: float, analysis_window: float = ANALYSIS_WINDOW_S) -> List[Tuple[float, float]]:
    """``auditok.split(bytes, min_dur, max_dur, max_silence, energy_threshold, drop_trailing_silence=True)`` given the per-block
    sums of squares of the region -> [(start, end)] seconds relative to the region."""
    if min_dur <= 0 or max_dur <= 0 or max_silence < 0:
        raise ValueError("min_dur, max_dur must be > 0 and max_silence >= 0")
    nb = len(sumsq)
    counts = np.full(nb, window, dtype=np.int64)
    if nb:
        counts[-1] = n_samples - (nb - 1) * window
    valid = energy_flags(np.asarray(sumsq), counts, energy_threshold)
    min_len, max_len = windows_for(min_dur, analysis_window, True), windows_for(max_dur, analysis_window, False)
    if min_len > max_len:
        raise ValueError("'min_dur' results in more analysis windows than 'max_dur'")
    toks = tokenize_flags(valid.tolist(), min_len, max_len, windows_for(max_silence, analysis_window, False), True)
    block_dur = window / sr
    out = []
    for first, blocks in toks:
        samples = blocks * window
        if first + blocks == nb:
            samples -= window - int(counts[-1])
        start = first * block_dur
        out.append((start, start + samples / sr))
    return out


def clock_split(start: float, end: float, chunk: float, min_duration: float) -> List[Tuple[float, float]]:
    """scene_detection_backends/utils.py:155-200 (brute_force_split)"""
    total = end - start
    if total <= 0:
        return []
    out = []
    for i in range(int(np.ceil(total / max(chunk, min_duration)))):
        a, b = start + i * chunk, min(start + (i + 1) * chunk, end)
        if b - a >= min_duration:
            out.append((a, b))
    return out


@dataclass
class Scene:
    """``SceneInfo`` (base.py:37-98) without the reference import."""
    start_sec: float
    end_sec: float
    scene_path: Optional[Path] = None
    detection_pass: int = 0
    metadata: Dict[str, Any] = field(default_factory=dict)

    @property
    def duration_sec(self) -> float:
        return self.end_sec - self.start_sec

    def to_legacy_tuple(self):
        return (self.scene_path, self.start_sec, self.end_sec, self.duration_sec)


EnergyFn = Callable[[Sequence[Tuple[int, int]], int], List[np.ndarray]]


FineFn = Callable[[List[Tuple[int, int, int]]], Dict[int, List[Tuple[float, float]]]]


def two_pass(cfg: SceneConfig, n_samples: int, sr: int, energy: EnergyFn, fine_split: Optional[FineFn] = None
             ) -> Tuple[List[Scene], List[Tuple[float, float]], Dict[str, int]]:
    """The driver of auditok_backend.py:229-524 on top of an energy provider ``energy(regions [(start, len)], window) -> per-region
    uint64 sums of squares`` (the GPU kernel in the product, a numpy twin in the CPU tests).  Two provider calls in total.
    ``fine_split([(story line, first sample, end sample)]) -> {story line: [(start, end)] seconds relative to it}`` replaces the
    energy pass 2 (the Silero-style detector, silero_backend.py:188-271)."""
    total = n_samples / sr
    window = int(ANALYSIS_WINDOW_S * sr)
    if window <= 0:
        raise ValueError("sample rate too low for a 50 ms analysis window")

    def clamp(s, e):
        s2 = max(0.0, s - cfg.pad_edges_s)
        e2 = min(total, e + cfg.pad_edges_s)
        return s2, max(e2, s2)

    counters = {"direct": 0, "granular": 0, "brute_force": 0}
    if n_samples == 0:
        return [], [], counters
    (ss1,) = energy([(0, n_samples)], window)
    story = split_region(ss1, n_samples, window, sr, cfg.pass1_min_duration, cfg.pass1_max_duration, min(total * 0.95, cfg.pass1_max_silence),
                         cfg.pass1_energy_threshold)
    big = [(k, int(rs * sr), int(re_ * sr)) for k, (rs, re_) in enumerate(story) if not (cfg.min_duration <= re_ - rs <= cfg.max_duration)]
    fine: Dict[int, List[Tuple[float, float]]] = {}
    if big and fine_split is not None:
        fine = dict(fine_split(big))
        for (k, _, _) in big:
            fine.setdefault(k, [])
    elif big:
        sums = energy([(a, b - a) for (_, a, b) in big], window)
        for (k, a, b), ss in zip(big, sums):
            dur = story[k][1] - story[k][0]
            fine[k] = split_region(ss, b - a, window, sr, cfg.pass2_min_duration, cfg.pass2_max_duration, min(dur * 0.95, cfg.pass2_max_silence),
                                   cfg.pass2_energy_threshold)
    scenes: List[Scene] = []
    for k, (rs, re_) in enumerate(story):
        if k not in fine:
            s, e = clamp(rs, re_)
            scenes.append(Scene(s, e, detection_pass=1))
            counters["direct"] += 1
        elif fine[k]:
            for (a, b) in fine[k]:
                if (rs + b) - (rs + a) < cfg.min_duration:
                    continue
                s, e = clamp(rs + a, rs + b)
                scenes.append(Scene(s, e, detection_pass=2))
                counters["granular"] += 1
        elif cfg.brute_force_fallback:
            for (a, b) in clock_split(rs, re_, cfg.brute_force_chunk_s, cfg.min_duration):
                s, e = clamp(a, b)
                scenes.append(Scene(s, e, detection_pass=2, metadata={"split_method": "brute_force"}))
                counters["brute_force"] += 1
    return scenes, story, counters


class B200SceneDetector:
    """SceneDetector protocol (base.py:185-250); ctor as ``AuditokSceneDetector`` (typed config or legacy kwargs)."""

    def __init__(self, config: Optional[SceneConfig] = None, device: str = "cuda", **kwargs: Any):
        self._config = config if config is not None else config_from_kwargs(kwargs)
        if self._config.assist_processing:
            logger.warning("b200-auditok: assist_processing (band-pass + DRC before pass 2) is not implemented; pass 2 runs on the raw audio")
        self._device = device
        self._lib = None
        self._last_result = None

    @property
    def name(self) -> str:
        return "b200-auditok"

    @property
    def display_name(self) -> str:
        return "B200 Auditok-style (Silence-Based, CUDA energy gate)"

    # -- device side ------------------------------------------------------------------------------------
    def _energy_provider(self, audio) -> EnergyFn:
        import torch
        from . import _lib
        if not torch.cuda.is_available():
            raise _lib.WjbError("B200SceneDetector needs a CUDA device; there is no CPU fallback")
        if self._lib is None:
            self._lib = _lib.load()
        dev = torch.device(self._device)
        if isinstance(audio, torch.Tensor):
            a = audio.to(device=dev, dtype=torch.float32).contiguous()
        else:
            a = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32)).to(dev, non_blocking=True)
        lib = self._lib

        def energy(regions, window):
            lens = [l for (_, l) in regions]
            nwin = [(l + window - 1) // window for l in lens]
            base = np.concatenate([[0], np.cumsum(nwin)]).astype(np.int64)
            meta = torch.tensor([[s for (s, _) in regions] + [0], lens + [0], base.tolist()], dtype=torch.int64).to(dev)
            out = torch.empty(int(base[-1]), dtype=torch.int64, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.wjb_scene_energy(_lib.ptr(a), a.numel(), _lib.ptr(meta[0]), _lib.ptr(meta[1]), _lib.ptr(meta[2]), len(regions),
                                                int(window), _lib.ptr(out), int(base[-1]), _lib.stream_ptr()), "wjb_scene_energy")
            host = out.cpu().numpy().view(np.uint64)
            return [host[base[i]: base[i + 1]] for i in range(len(regions))]

        return energy

    # -- arrays in memory -------------------------------------------------------------------------------
    def detect(self, audio, sample_rate: int = 16000) -> Tuple[List[Scene], List[Tuple[float, float]], Dict[str, int]]:
        """audio: float32 mono in [-1, 1] (numpy, or a torch tensor already on the device) -> (scenes, story lines, counters)."""
        n = int(audio.shape[0])
        return two_pass(self._config, n, int(sample_rate), self._energy_provider(audio))

    # -- protocol ---------------------------------------------------------------------------------------
    def detect_scenes(self, audio_path: Path, output_dir: Path, media_basename: str, **kwargs: Any):
        from .audioio import read_wav_mono, write_wav_pcm16
        try:
            from whisperjav.modules.scene_detection_backends.base import (SceneDetectionError, SceneDetectionResult,  # type: ignore
                                                                            SceneInfo)
        except Exception:  # WhisperJAV not importable: structural twins
            SceneDetectionError, SceneDetectionResult, SceneInfo = RuntimeError, _Result, Scene
        t0 = time.time()
        try:
            audio, sr = read_wav_mono(audio_path)
        except Exception as e:
            raise SceneDetectionError(f"Failed to load audio file {audio_path}: {e}") from e
        total = len(audio) / sr
        scenes, story, counters = self.detect(audio, sr)
        coarse = [{"scene_index": i, "start_time_seconds": round(a, 3), "end_time_seconds": round(b, 3), "duration_seconds": round(b - a, 3)}
                  for i, (a, b) in enumerate(story)]
        out = []
        if scenes:
            Path(output_dir).mkdir(parents=True, exist_ok=True)
        for i, sc in enumerate(scenes):
            a, b = int(sc.start_sec * sr), int(sc.end_sec * sr)
            if b <= a:
                raise ValueError(f"Empty audio data for scene {i}")  # save_scene_wav, utils.py:131-132
            path = Path(output_dir) / f"{media_basename}_scene_{i:04d}.wav"
            write_wav_pcm16(path, audio[a:b], sr)
            out.append(SceneInfo(start_sec=sc.start_sec, end_sec=sc.end_sec, scene_path=path, detection_pass=sc.detection_pass,
                                 metadata=dict(sc.metadata)))
        cfg = self._config
        params = {"max_duration": cfg.max_duration, "min_duration": cfg.min_duration, "pass1_max_silence": cfg.pass1_max_silence,
                  "pass1_energy_threshold": cfg.pass1_energy_threshold, "pass2_max_duration": cfg.pass2_max_duration,
                  "pass2_max_silence": cfg.pass2_max_silence, "pass2_energy_threshold": cfg.pass2_energy_threshold,
                  "assist_processing": False, "brute_force_fallback": cfg.brute_force_fallback, "counters": counters}
        self._last_result = SceneDetectionResult(scenes=out, method=self.name, audio_duration_sec=total, parameters=params,
                                                 processing_time_sec=time.time() - t0, coarse_boundaries=coarse)
        return self._last_result

    def cleanup(self) -> None:
        self._last_result = None


@dataclass
class SileroSceneConfig(SceneConfig):
    """Field for field ``SileroSceneConfig`` (silero_backend.py:27-49)."""
    silero_threshold: float = 0.06
    silero_neg_threshold: float = 0.15
    silero_min_silence_ms: int = 1500
    silero_min_speech_ms: int = 100
    silero_max_speech_s: float = 600.0
    silero_min_silence_at_max: int = 500
    silero_speech_pad_ms: int = 200

    def __post_init__(self):
        super().__post_init__()
        self.assist_processing = False


def silero_config_from_kwargs(kw: dict) -> SileroSceneConfig:
    """silero_backend.py:97-149: the auditok fields from the legacy kwargs, but ``max_duration`` defaults to 420 s, ``pass2_max_duration``
    is re-derived from it unless given, the brute-force chunk stays at 29 s."""
    base = config_from_kwargs(kw)
    user_p2 = "pass2_max_duration_s" in kw or "pass2_max_duration" in kw
    return SileroSceneConfig(
        max_duration=float(kw.get("max_duration_s", kw.get("max_duration", 420.0))), min_duration=base.min_duration,
        pass1_min_duration=base.pass1_min_duration, pass1_max_duration=base.pass1_max_duration, pass1_max_silence=base.pass1_max_silence,
        pass1_energy_threshold=base.pass1_energy_threshold, pass2_min_duration=base.pass2_min_duration,
        pass2_max_duration=base.pass2_max_duration if user_p2 else None, pass2_max_silence=base.pass2_max_silence,
        pass2_energy_threshold=base.pass2_energy_threshold, brute_force_fallback=base.brute_force_fallback,
        brute_force_chunk_s=float(kw.get("brute_force_chunk_s", 29.0)), pad_edges_s=base.pad_edges_s, verbose_summary=base.verbose_summary,
        force_mono=base.force_mono,
        silero_threshold=float(kw.get("silero_threshold", 0.06)), silero_neg_threshold=float(kw.get("silero_neg_threshold", 0.15)),
        silero_min_silence_ms=int(kw.get("silero_min_silence_ms", 1500)), silero_min_speech_ms=int(kw.get("silero_min_speech_ms", 100)),
        silero_max_speech_s=float(kw.get("silero_max_speech_s", 600.0)), silero_min_silence_at_max=int(kw.get("silero_min_silence_at_max", 500)),
        silero_speech_pad_ms=int(kw.get("silero_speech_pad_ms", 200)))


class B200SileroSceneDetector(B200SceneDetector):
    """``SileroSceneDetector`` (silero_backend.py:52-299): pass 1 is the energy gate of the parent, pass 2 runs the Silero-class VAD
    over every oversized chapter -- here all of them in ONE batch of the GPU gate (``vad.VadB200.probs``) -- and turns the per-window
    probabilities into speech regions with the hysteresis of ``get_speech_timestamps`` (``hostlogic.probs_to_regions`` with the
    explicit ``neg_threshold`` the reference passes; seconds rounded to 0.1 s as silero-vad's ``return_seconds=True`` does).  A chapter
    in which the VAD finds nothing falls through to the brute-force split, as in the reference.  ``silero_min_silence_at_max`` is
    accepted; the split of a region longer than ``silero_max_speech_s`` (600 s by default) cuts at the limit.  The VAD network itself is
    the Silero-class stack of ``vad.py`` (weights not available offline -- unpinned)."""

    def __init__(self, config: Optional[SileroSceneConfig] = None, device: str = "cuda", vad=None, **kwargs: Any):
        cfg = config if config is not None else silero_config_from_kwargs(kwargs)
        super().__init__(config=cfg, device=device)
        self._silero_config = cfg
        self._vad = vad               # anything with ``probs(audio [B, S] device fp32, n_samples [B] device int32) -> [B, windows]``
        self._vad_segments: List[Dict[str, float]] = []

    @property
    def name(self) -> str:
        return "b200-silero"

    @property
    def display_name(self) -> str:
        return "B200 Silero-style (energy pass 1 + CUDA VAD pass 2)"

    def _region_probs(self, clips: List[np.ndarray]) -> List[np.ndarray]:
        import torch
        from .vad import WINDOW, VadB200
        if self._vad is None:
            self._vad = VadB200(device=self._device)
        S = max(max(len(c) for c in clips), WINDOW)
        host = torch.zeros(len(clips), S, dtype=torch.float32)
        for i, c in enumerate(clips):
            host[i, : len(c)] = torch.from_numpy(np.ascontiguousarray(c, dtype=np.float32))
        ns = torch.tensor([len(c) for c in clips], dtype=torch.int32)
        dev = getattr(self._vad, "device", "cpu")
        probs = self._vad.probs(host.to(dev), ns.to(dev)).cpu().numpy()
        return [probs[i, : (len(c) + WINDOW - 1) // WINDOW] for i, c in enumerate(clips)]

    def _fine_split(self, audio: np.ndarray, sr: int) -> FineFn:
        from . import hostlogic as H
        from .vad import WINDOW
        cfg = self._silero_config

        def fine(big):
            clips = []
            for (_, a, b) in big:
                x = np.asarray(audio[a:b], dtype=np.float32)
                if sr != 16000:  # nearest-index decimation (backends/silero.py:411-414); the hot path hands 16 kHz audio
                    x = x[np.linspace(0, len(x) - 1, int(len(x) * 16000 / sr)).astype(int)] if len(x) else x
                clips.append(x)
            out: Dict[int, List[Tuple[float, float]]] = {}
            for (k, a, _), c, p in zip(big, clips, self._region_probs(clips) if clips else []):
                dur = len(c) / 16000.0
                regs = H.probs_to_regions(p, dur, frame_ms=1000.0 * WINDOW / 16000, threshold=cfg.silero_threshold,
                                          neg_threshold=cfg.silero_neg_threshold, min_speech_duration_ms=cfg.silero_min_speech_ms,
                                          min_silence_duration_ms=cfg.silero_min_silence_ms, speech_pad_ms=cfg.silero_speech_pad_ms,
                                          max_speech_duration_s=cfg.silero_max_speech_s)
                subs = [(max(round(r.start_sample / 16000.0, 1), 0), min(round(r.end_sample / 16000.0, 1), dur)) for r in regs]
                out[k] = subs
                t0 = a / sr
                self._vad_segments.extend({"start_sec": round(t0 + s0, 3), "end_sec": round(t0 + s1, 3)} for s0, s1 in subs)
            return out

        return fine

    def detect(self, audio, sample_rate: int = 16000):
        import torch
        self._vad_segments = []
        host = audio.detach().cpu().numpy() if isinstance(audio, torch.Tensor) else np.asarray(audio, dtype=np.float32)
        return two_pass(self._config, int(host.shape[0]), int(sample_rate), self._energy_provider(audio), self._fine_split(host, int(sample_rate)))

    def detect_scenes(self, audio_path: Path, output_dir: Path, media_basename: str, **kwargs: Any):
        res = super().detect_scenes(audio_path, output_dir, media_basename, **kwargs)
        cfg = self._silero_config
        res.parameters.update({"silero_threshold": cfg.silero_threshold, "silero_neg_threshold": cfg.silero_neg_threshold,
                               "silero_min_silence_ms": cfg.silero_min_silence_ms, "silero_min_speech_ms": cfg.silero_min_speech_ms,
                               "silero_max_speech_s": cfg.silero_max_speech_s, "silero_speech_pad_ms": cfg.silero_speech_pad_ms})
        if hasattr(res, "vad_segments"):
            res.vad_segments = self._vad_segments or None
        return res

    def cleanup(self) -> None:
        super().cleanup()
        self._vad = None
        self._vad_segments = []


@dataclass
class _Result:
    """``SceneDetectionResult`` (base.py:100-183) when WhisperJAV is not importable."""
    scenes: List[Scene]
    method: str
    audio_duration_sec: float
    parameters: Dict[str, Any] = field(default_factory=dict)
    processing_time_sec: float = 0.0
    coarse_boundaries: Optional[List[Dict[str, Any]]] = None
    vad_segments: Optional[List[Dict[str, Any]]] = None

    @property
    def num_scenes(self) -> int:
        return len(self.scenes)

    def to_legacy_tuples(self):
        return [s.to_legacy_tuple() for s in self.scenes]
