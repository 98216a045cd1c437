"""Timeline of one decoder step from the GEMM trace hook: per-GEMM lifetime of CTA 0 and the gaps between GEMMs
(where LayerNorm / attention / sampling run), large-v3, batch 64.  Debugging aid; numbers are global-timer based."""
import os
import sys
from collections import defaultdict
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from whisperjav_b200 import _lib, model as M  # noqa: E402

lib = _lib.load()
m = M.load_model("large-v3", max_batch=64)
xa = torch.randn(64, 1500, 1280, device="cuda", dtype=torch.float16)
trace = torch.zeros(32 + 512 * 32, dtype=torch.int64, device="cuda")
lib.wjb_debug_gemm_trace(_lib.ptr(trace))
m.decode_features(xa, without_timestamps=True, sample_len=14)
torch.cuda.synchronize()
lib.wjb_debug_gemm_trace(None)
t = trace.cpu()
rows = t[32:].view(512, 32)
rows = rows[rows[:, 29].argsort()]
idx = rows[:, 29].tolist()
N = (rows[:, 28] >> 32).tolist()
K = (rows[:, 28] & 0xffffffff).tolist()
gt0 = rows[:, 1].tolist()
gt8 = rows[:, 17].tolist()
ck = rows[:, 0:18:2].double()
# a step = the launches between two logits GEMMs
ends = [i for i, n in enumerate(N) if n > 50000]
a, b = ends[-2] + 1, ends[-1] + 1
print(f"step spans launches {idx[a]}..{idx[b - 1]} ({b - a} GEMMs), wall {(gt8[b - 1] - gt8[a - 1]) / 1000:.1f} us")
names = {(3840, 1280): "qkv", (5120, 1280): "fc1", (1280, 5120): "fc2"}
order = ["qkv", "out", "cq", "cout", "fc1", "fc2"]
dur = defaultdict(float)
gap = defaultdict(float)
kloop = defaultdict(float)
cnt = defaultdict(int)
pos = 0
for i in range(a, b):
    if N[i] > 50000:
        nm = "logits"
    else:
        nm = order[pos % 6]
        pos += 1
    d = gt8[i] - gt0[i]
    g = gt0[i] - gt8[i - 1]
    dur[nm] += d
    gap[nm] += g
    kloop[nm] += (ck[i, 5] - ck[i, 4]).item() / 1.9
    cnt[nm] += 1
    if 6 * 10 <= i - a < 6 * 11:
        rel = ((ck[i] - ck[i, 0]) / 1.9).tolist()
        print(f"  layer10 {nm:6s} N={N[i]} K={K[i]} gap_before={g:6.0f} life={d:6.0f} ns | pro={rel[1]:.0f} wait={rel[2]:.0f} tma0={rel[3]:.0f} "
              f"full0={rel[4]:.0f} mma_done={rel[5]:.0f} stored={rel[7]:.0f} exit={rel[8]:.0f}")
print("per step totals (us): name count life gap_before k-loop")
tl = tg = 0
for nm in order + ["logits"]:
    print(f"  {nm:6s} {cnt[nm]:3d} life {dur[nm] / 1000:7.1f} gap_before {gap[nm] / 1000:7.1f} kloop {kloop[nm] / 1000:7.1f}")
    tl += dur[nm]
    tg += gap[nm]
print(f"  sum life {tl / 1000:.1f} us, sum gaps {tg / 1000:.1f} us")
