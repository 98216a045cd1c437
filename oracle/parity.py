"""Parity checker shared by tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s ``parity_check`` block: CUDA path (through the
C-ABI, via ``whisperjav_b200.model.WhisperB200``) against the CPU oracle on the same weights and inputs.

TEST INFRASTRUCTURE ONLY -- nothing under ``whisperjav_b200/`` may import this module (or anything else in ``oracle/``).

How token identity is judged (north_star: "decoded token ids identical under greedy"):

* the device decodes freely and records the raw fp16 logits of every step (``wjb_decode_set_trace``);
* the oracle is then *teacher-forced along the device's own token sequence* (``whisper_oracle.decode(forced_tokens=...)``), so both
  sides evaluate the decoder on exactly the same prefix at every step.  Per step this yields
    - ``dlogit``: max |logit_gpu - logit_oracle| over the vocabulary (raw logits, before the filters), expressed in fp16 quanta
      of the step's top logit -- the kernel-correctness measure, independent of how close the top-2 candidates are;
    - whether the device's token is the oracle's own arg-max on that prefix.  If it is at every step, the oracle's free-running
      greedy decode is, by induction, the same sequence: the window is *token-identical*.
    - otherwise the step is a tie-break: it is accepted only if the device's token sits within ``tie_quanta`` fp16 quanta of the
      oracle's top filtered logit (logits are fp16 numbers: upstream ``(x @ emb.T).float()`` under ``fp16=True``), and it is counted.
  Nothing is left unchecked after a tie-break: the comparison continues on the device's branch to the end of the sequence.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import whisper_oracle as wo


def fp16_quantum(x: float) -> float:
    """Spacing of fp16 numbers at magnitude |x| (normal range)."""
    x = abs(float(x))
    if x < 2.0 ** -14:
        return 2.0 ** -24
    return 2.0 ** (math.floor(math.log2(x)) - 10)


def decode_parity(model, weights, dims, xa: torch.Tensor, *, tie_quanta: float = 4.0, logit_quanta: float = 16.0,
                  logit_rms_quanta: float = 3.0, logprob_tol_per_step: float = 0.02, prepared=None, **decode_kw) -> Dict:
    """Greedy decode of encoder output ``xa`` (device tensor) on the GPU vs the oracle.  ``decode_kw`` are DecodingOptions
    fields understood by both sides (language, without_timestamps, max_initial_timestamp, sample_len, suppress_tokens ...).
    Returns a report dict; ``report["ok"]`` is the verdict, ``report["failures"]`` says why not."""
    res, trace = model.decode_trace(xa, **decode_kw)
    n_initial = trace["n_initial"]
    gl = trace["logits"]  # [steps][B][V]
    pw = prepared if prepared is not None else wo.prepare_weights(weights, True)
    opts = wo.DecodingOptions(**decode_kw)
    ref, rl = wo.decode(pw, dims, None, opts, True, audio_features=xa.float().cpu(), return_logits=True,
                        forced_tokens=[r.tokens for r in res])
    B = xa.shape[0]
    rows, failures = [], []
    all_dq, all_margin, all_rms = [], [], []
    for b in range(B):
        toks = res[b].tokens
        n_steps = min(len(toks) + 1, len(rl))  # the step that produced EOT counts (unless sample_len ran out first)
        n_steps = min(n_steps, gl.shape[0] - (n_initial - 1))
        dq, drms, ties, bad = [], [], [], []
        for i in range(n_steps):
            g = gl[n_initial - 1 + i, b]
            r = rl[i][b]
            q = fp16_quantum(float(r.max()))
            d = float((g - r).abs().max()) / q
            rms = float((g - r).pow(2).mean().sqrt()) / q
            dq.append(d)
            drms.append(rms)
            # two numbers per step: the largest difference over the 51 k vocabulary entries (a Gaussian extreme, ~4.5 x the
            # rms) and the rms itself; rounding differences alone give rms ~ 1 quantum (scripts/synth_chaos.py)
            if d > logit_quanta or rms > logit_rms_quanta:
                bad.append({"step": i, "dlogit_quanta": d, "rms_quanta": rms})
        picks, gaps = ref[b].picks, ref[b].forced_gap
        fed = list(toks) + [opts_eot(dims)]
        for i in range(min(len(picks), n_steps)):
            if picks[i] != fed[i]:
                q = fp16_quantum(float(rl[i][b].max()))
                ties.append({"step": i, "gpu": int(fed[i]), "oracle": int(picks[i]), "gap_quanta": gaps[i] / q})
        row = {"b": b, "len": len(toks), "steps_checked": n_steps, "identical": not ties, "tie_breaks": ties,
               "dlogit_quanta_max": max(dq) if dq else 0.0, "dlogit_quanta_median": float(np.median(dq)) if dq else 0.0,
               "dlogit_rms_quanta_max": max(drms) if drms else 0.0,
               "oracle_margin_min": min(ref[b].margins) if ref[b].margins else None,
               "sum_logprob_gpu": res[b].sum_logprob, "sum_logprob_oracle": ref[b].sum_logprob,
               "no_speech_gpu": res[b].no_speech_prob, "no_speech_oracle": ref[b].no_speech_prob}
        rows.append(row)
        all_rms += drms
        all_dq += dq
        all_margin += list(ref[b].margins)
        if bad:
            failures.append({"b": b, "why": "step logits differ from the oracle on the same prefix", "steps": bad[:4]})
        for t in ties:
            if t["gap_quanta"] > tie_quanta:
                failures.append({"b": b, "why": "device token is not a near-tie of the oracle's arg-max", **t})
        if abs(res[b].sum_logprob - ref[b].sum_logprob) > logprob_tol_per_step * max(1, n_steps):
            failures.append({"b": b, "why": "sum_logprob", "gpu": res[b].sum_logprob, "oracle": ref[b].sum_logprob})
        if abs(res[b].no_speech_prob - ref[b].no_speech_prob) > 1e-3 + 0.03 * ref[b].no_speech_prob:
            failures.append({"b": b, "why": "no_speech_prob", "gpu": res[b].no_speech_prob, "oracle": ref[b].no_speech_prob})
    n_steps_total = sum(r["steps_checked"] for r in rows)
    n_ties = sum(len(r["tie_breaks"]) for r in rows)
    m = np.asarray(all_margin) if all_margin else np.zeros(1)
    return {"ok": not failures, "failures": failures, "windows": B, "identical_windows": sum(r["identical"] for r in rows),
            "steps_checked": n_steps_total, "tie_breaks": n_ties, "dlogit_quanta_max": max(all_dq) if all_dq else 0.0,
            "dlogit_quanta_p99": float(np.quantile(all_dq, 0.99)) if all_dq else 0.0,
            "dlogit_rms_quanta_max": max(all_rms) if all_rms else 0.0,
            "oracle_margin_median": float(np.median(m)), "oracle_margin_frac_below_0.1": float((m < 0.1).mean()),
            "tolerances": {"tie_quanta": tie_quanta, "logit_quanta": logit_quanta, "logit_rms_quanta": logit_rms_quanta}, "rows": rows, "tokens": [r.tokens for r in res]}


def opts_eot(dims) -> int:
    return wo.SpecialTokens(dims.n_vocab, language="en").eot


def gpu_mel(model, clips: Sequence[np.ndarray], n_frames: int = 3000, reflect_total: int = 0) -> torch.Tensor:
    """Clips -> the device's time-major log-mel [B, n_frames + 2, n_mels] (fp16)."""
    S = max(len(c) for c in clips)
    audio = torch.zeros(len(clips), S)
    for i, c in enumerate(clips):
        audio[i, : len(c)] = torch.from_numpy(np.asarray(c, dtype=np.float32))
    ns = torch.tensor([len(c) for c in clips], dtype=torch.int32)
    return model.log_mel(audio.to(model.device), ns.to(model.device), n_frames=n_frames, layout="time", reflect_total=reflect_total)


def oracle_mel_windows(clips: Sequence[np.ndarray], dims) -> torch.Tensor:
    """What upstream transcribe() feeds the encoder for the first window of every clip: [B, n_mels, 3000] fp32."""
    return torch.stack([wo.pad_or_trim(wo.log_mel_spectrogram(a, dims.n_mels, padding=wo.N_SAMPLES)[:, : len(a) // 160], wo.N_FRAMES)
                        for a in clips])


def encoder_parity(model, weights, dims, mel_tm: torch.Tensor, tap_every: int = 0, prepared=None) -> Dict:
    """GPU encoder vs the oracle fed the *same* fp16 mel.  north_star tolerance: hidden states within 1e-2 relative."""
    pw = prepared if prepared is not None else wo.prepare_weights(weights, True)
    mel_in = mel_tm[:, 1:-1].permute(0, 2, 1).float().cpu()
    if tap_every:
        xa, taps = model.encode(mel_tm, tap_every=tap_every)
        ref, layers = wo.encoder_forward(pw, dims, mel_in, True, return_layers=True)
    else:
        xa, taps = model.encode(mel_tm), None
        ref, layers = wo.encoder_forward(pw, dims, mel_in, True), None
    got = xa.float().cpu()
    out = {"rel_fro": float((got - ref).norm() / ref.norm()), "max_abs": float((got - ref).abs().max()), "ref_absmax": float(ref.abs().max()),
           "per_row_rel_max": float(((got - ref).norm(dim=-1) / ref.norm(dim=-1).clamp_min(1e-6)).max())}
    if taps is not None:
        out["taps"] = []
        for k in range(taps.shape[0]):
            r = layers[(k + 1) * tap_every - 1]
            t = taps[k].float().cpu()
            out["taps"].append({"after_block": (k + 1) * tap_every, "rel_fro": float((t - r).norm() / r.norm())})
    out["ok"] = out["rel_fro"] <= 1e-2 and all(t["rel_fro"] <= 1e-2 for t in out.get("taps", []))
    return out, xa


def transcribe_parity(model, weights, dims, clips: Sequence[np.ndarray], *, tie_quanta: float = 16.0, prepared=None, **kw) -> Dict:
    """End to end (device mel + encoder + decoder + seek loop vs the oracle's) with nothing left unchecked after a divergence:
    the oracle *follows the device window by window*.  For every window the device decoded (``_record_windows``) the oracle builds
    that window from its own log-mel, runs its own encoder, is teacher-forced along the device's tokens and must find every one
    of them to be its arg-max or within ``tie_quanta`` fp16 quanta of it (the two encoders differ by ~1e-3 relative, which moves
    logits by a few quanta: larger than in the decoder-only comparison).  Then the host logic is replayed on the device's own
    decode result: the oracle's ``slice_segments`` (and, with ``word_timestamps``, its ``add_word_timestamps`` on its own encoder
    output) must give the device's segments, words and the next seek (word mode: within one DTW frame = 2 mel frames; the replay
    then continues from the device's seek).  Clips whose every token is the oracle's arg-max are, by induction, token-identical to
    the oracle's free-running transcribe()."""
    from . import timing_oracle as to
    pw = prepared if prepared is not None else wo.prepare_weights(weights, True)
    dec_keys = {"language", "task", "without_timestamps", "max_initial_timestamp", "suppress_tokens", "suppress_blank", "sample_len"}
    words_mode = bool(kw.get("word_timestamps"))
    got = model.transcribe_batch(list(clips), _record_windows=True, **kw)
    opts = wo.DecodingOptions(**{k: v for k, v in kw.items() if k in dec_keys})
    language, task = kw.get("language", "ja"), kw.get("task", "transcribe")
    tok = wo.SpecialTokens(dims.n_vocab, language=language, task=task)
    failures, rows = [], []
    for ci, (a, g) in enumerate(zip(clips, got)):
        mel = wo.log_mel_spectrogram(a, dims.n_mels, padding=wo.N_SAMPLES)
        content = mel.shape[-1] - wo.N_FRAMES
        seek, ties, steps, n_words, words_close = 0, 0, 0, 0, 0
        last_speech = 0.0
        dev_segments = g["segments"]
        for wi, w in enumerate(g["windows"]):
            if abs(w["seek"] - seek) > (2 if words_mode else 0):
                failures.append({"clip": ci, "why": "seek sequence", "device": w["seek"], "replayed": seek})
                break
            seek = w["seek"]
            size = min(wo.N_FRAMES, content - seek)
            win = wo.pad_or_trim(mel[:, seek: seek + size], wo.N_FRAMES)
            xa_o = wo.encoder_forward(pw, dims, win[None], True)
            ref = wo.decode(pw, dims, None, opts, True, audio_features=xa_o, forced_tokens=[w["tokens"]])[0]
            fed = list(w["tokens"]) + [tok.eot]
            for i, (pick, gap) in enumerate(zip(ref.picks, ref.forced_gap)):
                steps += 1
                if pick != fed[i]:
                    ties += 1
                    q = fp16_quantum(32.0)
                    if not gap <= tie_quanta * q:
                        failures.append({"clip": ci, "seek": seek, "step": i, "why": "device token is not a near-tie of the oracle's arg-max",
                                         "gap_quanta": gap / q})
            # teacher-forced scores on two slightly different encoder outputs: a few quanta per token
            if abs(ref.avg_logprob - w["avg_logprob"]) > 0.06 or abs(ref.no_speech_prob - w["no_speech_prob"]) > 2e-3 + 0.05 * ref.no_speech_prob:
                failures.append({"clip": ci, "seek": seek, "why": "avg_logprob / no_speech_prob", "device": [w["avg_logprob"], w["no_speech_prob"]],
                                 "oracle": [ref.avg_logprob, ref.no_speech_prob]})
            # host logic replay (transcribe.py's no-speech skip, slicing, word timestamps, seek advance) on the device's decode result
            nst, lpt = kw.get("no_speech_threshold", 0.6), kw.get("logprob_threshold", -1.0)
            skip = nst is not None and w["no_speech_prob"] > nst and not (lpt is not None and w["avg_logprob"] > lpt)
            if skip:
                seek += size
                continue
            fields = {"temperature": w["temperature"], "avg_logprob": w["avg_logprob"], "no_speech_prob": w["no_speech_prob"]}
            info: dict = {}
            cur, adv = wo.slice_segments(w["tokens"], tok, seek, size, fields, clear=False, info=info)
            time_offset = float(seek * wo.HOP_LENGTH / wo.SAMPLE_RATE)
            seek += adv
            if words_mode:
                to.add_word_timestamps(pw, dims, cur, xa_o, size, language=language, task=task, last_speech_timestamp=last_speech)
                if not info["single_timestamp_ending"]:
                    lwe = to.get_end(cur)
                    if lwe is not None and lwe > time_offset:
                        seek = round(lwe * wo.FRAMES_PER_SECOND)
                lwe = to.get_end(cur)
                if lwe is not None:
                    last_speech = lwe
            wo.clear_empty_segments(cur)
            mine = [x for x in dev_segments if x["seek"] == w["seek"]]
            tol = 0.021 if words_mode else 1e-9
            if len(mine) != len(cur) or any(x["tokens"] != y["tokens"] or abs(x["start"] - y["start"]) > tol or abs(x["end"] - y["end"]) > tol
                                            for x, y in zip(mine, cur)):
                failures.append({"clip": ci, "seek": w["seek"], "why": "segments differ from the oracle's slicing of the same tokens",
                                 "device": [(x["start"], x["end"], len(x["tokens"])) for x in mine],
                                 "oracle": [(y["start"], y["end"], len(y["tokens"])) for y in cur]})
            elif words_mode:
                for x, y in zip(mine, cur):
                    if len(x["words"]) != len(y["words"]):
                        failures.append({"clip": ci, "seek": w["seek"], "why": "word count", "device": len(x["words"]), "oracle": len(y["words"])})
                        continue
                    for u, v in zip(x["words"], y["words"]):
                        n_words += 1
                        words_close += abs(u["start"] - v["start"]) <= 0.021 and abs(u["end"] - v["end"]) <= 0.021
        if not any(f["clip"] == ci for f in failures):
            if abs(seek - content) > (2 if words_mode else 0) and seek < content:
                failures.append({"clip": ci, "why": "device stopped before the end of the clip", "seek": seek, "content": content})
        rows.append({"clip": ci, "windows": len(g["windows"]), "steps": steps, "tie_breaks": ties, "identical": ties == 0, "segments": len(dev_segments),
                     "words": n_words, "words_within_1_frame": words_close})
    n_words = sum(r["words"] for r in rows)
    if words_mode and n_words and sum(r["words_within_1_frame"] for r in rows) < 0.9 * n_words:
        failures.append({"why": "fewer than 90 % of the word boundaries within one frame of the oracle's", "rows": rows})
    return {"ok": not failures, "failures": failures, "clips": len(clips), "identical_clips": sum(r["identical"] for r in rows),
            "steps_checked": sum(r["steps"] for r in rows), "tie_breaks": sum(r["tie_breaks"] for r in rows), "words": n_words,
            "words_within_1_frame": sum(r["words_within_1_frame"] for r in rows), "rows": rows, "tolerances": {"tie_quanta": tie_quanta}}
