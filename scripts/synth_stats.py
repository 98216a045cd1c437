"""Statistics of the synthetic-weight generator seen through the CPU oracle (no GPU): per-step top-2 margins of the filtered
logits, sequence lengths, token diversity, and the sensitivity of the logits to fp16 rounding (teacher-forced fp16-sim vs
fp32 run on the same tokens).  Used to tune whisperjav_b200/synth.py; the thresholds it reports are asserted in
tests/test_oracle_synth.py.

    python scripts/synth_stats.py [--model tiny] [--clips 4] [--seed 7] [--ts 0|1] [--sample-len N] [k=v generator overrides]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import whisper_oracle as wo  # noqa: E402
from whisperjav_b200.synth import DIMS, speech_shaped_audio, synth_weights  # noqa: E402


def stats(model="tiny", clips=4, seed=7, ts=1, sample_len=None, fp32_check=True, durations=(30.0, 12.0, 5.0, 21.7), **gen):
    dims = DIMS[model]
    w = synth_weights(dims, seed=seed, **gen)
    audio = [speech_shaped_audio(durations[i % len(durations)], 1000 + i) for i in range(clips)]
    mel = torch.stack([wo.pad_or_trim(wo.log_mel_spectrogram(a, dims.n_mels, padding=wo.N_SAMPLES)[:, : len(a) // 160], wo.N_FRAMES) for a in audio])
    t0 = time.time()
    pw = wo.prepare_weights(w, True)
    xa = wo.encoder_forward(pw, dims, mel, True)
    opts = wo.DecodingOptions(language="ja", without_timestamps=not ts, max_initial_timestamp=0.0, sample_len=sample_len)
    res, logits = wo.decode(pw, dims, None, opts, True, audio_features=xa, return_logits=True)
    m = np.concatenate([np.asarray(r.margins) for r in res])
    out = {"model": model, "gen": gen, "lens": [len(r.tokens) for r in res], "steps": int(m.size),
           "margin_median": float(np.median(m)), "margin_p02": float(np.quantile(m, 0.02)), "frac_below_0.1": float((m < 0.1).mean()),
           "frac_below_0.03": float((m < 0.03).mean()), "unique_tokens": len({t for r in res for t in r.tokens}),
           "max_repeat": max((sum(1 for a, b in zip(r.tokens, r.tokens[1:]) if a == b) for r in res), default=0),
           "top_logit_median": float(np.median([float(l.max()) for l in logits])),
           "avg_logprob": [round(r.avg_logprob, 3) for r in res], "no_speech": [round(r.no_speech_prob, 4) for r in res],
           "compression_ratio": [round(r.compression_ratio, 2) for r in res]}
    if fp32_check:
        # rounding sensitivity: the same tokens through the fp32 path; |dlogit| relative to the logit spread
        pw32 = wo.prepare_weights(w, False)
        xa32 = wo.encoder_forward(pw32, dims, mel, False)
        out["enc_rel_fp16_vs_fp32"] = float(((xa - xa32).norm() / xa32.norm()))
        _, l32 = wo.decode(pw32, dims, None, opts, False, audio_features=xa, return_logits=True, forced_tokens=[r.tokens for r in res])
        n = min(len(l32), len(logits))
        d = [float((logits[i] - l32[i]).abs().max()) for i in range(n)]
        out["dlogit_fp16_vs_fp32_max"] = max(d)
        out["dlogit_fp16_vs_fp32_median"] = float(np.median(d))
        out["logit_std"] = float(np.median([float(l.std()) for l in logits]))
    out["seconds"] = round(time.time() - t0, 1)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="tiny")
    ap.add_argument("--clips", type=int, default=4)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--ts", type=int, default=1)
    ap.add_argument("--sample-len", type=int, default=None)
    ap.add_argument("--no-fp32", action="store_true")
    ap.add_argument("gen", nargs="*")
    a = ap.parse_args()
    gen = {k: float(v) for k, v in (g.split("=") for g in a.gen)}
    print(json.dumps(stats(a.model, a.clips, a.seed, a.ts, a.sample_len, not a.no_fp32, **gen), indent=1))
