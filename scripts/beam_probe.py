"""Beam-search decode cost (large-v3, 64 windows): ms per decoder step for beam_size 1..3 next to the greedy loop."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from whisperjav_b200 import model as M  # noqa: E402

m = M.load_model("large-v3", max_batch=64)
xa = torch.randn(64, 1500, 1280, device="cuda", dtype=torch.float16)
for name, kw in [("greedy", {}), ("beam 2 (patience 1.2)", {"beam_size": 2, "patience": 1.2}), ("beam 3 (patience 1.5)", {"beam_size": 3, "patience": 1.5})]:
    m.decode_features(xa, without_timestamps=True, sample_len=24, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s0 = m.stats["decode_steps"]
    m.decode_features(xa, without_timestamps=True, sample_len=24, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    steps = m.stats["decode_steps"] - s0
    print(f"{name}: {steps} steps, {dt * 1e3 / steps:.3f} ms/step (incl. cross-K/V projection and graph capture)", flush=True)
