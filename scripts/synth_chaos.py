"""How strongly does the synthetic decoder amplify a perturbation of fp16-rounding size?  (CPU oracle only.)  The oracle is run
twice along the same token sequence, the second time with its encoder output perturbed by relative noise EPS; the per-step
max |dlogit| (in fp16 quanta of the top logit) is what a GPU-vs-oracle comparison must expect from rounding differences alone.

    python scripts/synth_chaos.py [--model large-v3] [--steps 48] [--eps 3e-4] [k=v generator overrides]"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import parity as P  # noqa: E402
from oracle import whisper_oracle as wo  # noqa: E402
from whisperjav_b200.synth import DIMS, speech_shaped_audio, synth_preset, synth_weights  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="large-v3")
ap.add_argument("--steps", type=int, default=48)
ap.add_argument("--eps", type=float, default=3e-4)
ap.add_argument("--ts", type=int, default=0)
ap.add_argument("gen", nargs="*")
a = ap.parse_args()
kw = synth_preset(a.model)
kw.update({k: float(v) for k, v in (g.split("=") for g in a.gen)})
kw["seed"] = int(kw["seed"])
dims = DIMS[a.model]
pw = wo.prepare_weights(synth_weights(dims, **kw), True)
clip = speech_shaped_audio(30.0, 2002)
mel = P.oracle_mel_windows([clip], dims)
xa = wo.encoder_forward(pw, dims, mel, True)
opts = wo.DecodingOptions(language="ja", without_timestamps=not a.ts, max_initial_timestamp=0.0, sample_len=a.steps)
res, l0 = wo.decode(pw, dims, None, opts, True, audio_features=xa, return_logits=True)
g = torch.Generator().manual_seed(1)
xa2 = (xa * (1 + a.eps * torch.randn(xa.shape, generator=g))).half().float()
_, l1 = wo.decode(pw, dims, None, opts, True, audio_features=xa2, return_logits=True, forced_tokens=[res[0].tokens])
dq = [float((x[0] - y[0]).abs().max()) / P.fp16_quantum(float(x[0].max())) for x, y in zip(l0, l1)]
print(json.dumps({"gen": kw, "len": len(res[0].tokens), "dq_first": dq[:3], "dq_max": max(dq), "dq_median": float(np.median(dq)), "dq_last": dq[-3:],
                  "max_repeat": sum(1 for x, y in zip(res[0].tokens, res[0].tokens[1:]) if x == y), "unique": len(set(res[0].tokens))}))
