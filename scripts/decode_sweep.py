"""Decode-step time vs. number of concurrent batch slices and GEMM variant (large-v3, B=64, 24 sampled tokens)."""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
code = r'''
import sys, time, torch, os
sys.path.insert(0, %r)
from whisperjav_b200 import model as M
print("WJB_DECODE_MEGA =", os.environ.get("WJB_DECODE_MEGA"), flush=True)
m = M.load_model("large-v3", max_batch=64)
xa = torch.randn(64, 1500, 1280, device="cuda", dtype=torch.float16)
for split, tc in [(1,1)]:
    os.environ["WJB_DECODE_SPLIT"] = str(split)
    m.decode_features(xa, without_timestamps=True, sample_len=24)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s0 = m.stats["decode_steps"]
    r = m.decode_features(xa, without_timestamps=True, sample_len=24)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    steps = m.stats["decode_steps"] - s0
    print("split", split, "tc_gemm", tc, "steps", steps, "ms/step %%.3f" %% (dt * 1e3 / steps), "(incl. cross-kv proj)", flush=True)
''' % str(ROOT)
subprocess.run([sys.executable, "-c", code], check=False)
