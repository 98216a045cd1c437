"""Timeline of CTA 0 across dependent decode-step GEMM launches inside a CUDA graph (debugging aid, wjb_debug_gemm_trace)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from whisperjav_b200 import _lib  # noqa: E402

lib = _lib.load()
DEV = "cuda"
M, REPS = 64, 40
NAMES = ["entry", "prologue", "pdl_wait", "tma0", "full0", "mma_done", "acc_ready", "stored", "exit"]


def run(N, K, bn, with_ln, pdl=0, S=None):
    lib.wjb_debug_set_pdl(pdl)
    Ws = [torch.randn(N, K, device=DEV, dtype=torch.float16) * 0.03 for _ in range(8)]
    A = torch.randn(M, K, device=DEV, dtype=torch.float16)
    b = torch.randn(N, device=DEV, dtype=torch.float16)
    out = torch.zeros(M, N, device=DEV, dtype=torch.float16)
    x = torch.randn(M, 1280, device=DEV, dtype=torch.float16)
    gm = torch.ones(1280, device=DEV, dtype=torch.float16)
    o = torch.empty_like(x)
    trace = torch.zeros(32 + 512 * 32, dtype=torch.int64, device=DEV)
    i = [0]

    def f():
        W = Ws[i[0] % 8]
        i[0] += 1
        if with_ln:
            _lib.check(lib.wjb_layernorm_f16(_lib.ptr(x), _lib.ptr(gm), _lib.ptr(gm), _lib.ptr(o), M, 1280, _lib.stream_ptr()), "ln")
        if S is None:
            _lib.check(lib.wjb_gemm_f16(_lib.ptr(A), K, 0, M, 1, K, _lib.ptr(W), N, K, _lib.ptr(b), None, _lib.ptr(out), N, 0, 0, bn, _lib.stream_ptr()), "tc")
        else:
            _lib.check(lib.wjb_gemm_step_f16(_lib.ptr(A), K, M, K, _lib.ptr(W), N, K, _lib.ptr(b), None, _lib.ptr(out), N, 0, bn, S, 1,
                                             _lib.stream_ptr()), "step")

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        lib.wjb_debug_gemm_trace(_lib.ptr(trace))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(REPS):
                f()
        lib.wjb_debug_gemm_trace(None)
        g.replay()
        torch.cuda.synchronize()
        trace.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
    t = trace.cpu()
    n = int(t[0])
    n = min(n, 512)
    rows = t[32:32 + n * 32].view(n, 32)
    print(f"--- N={N} K={K} bn={bn} S={S} ln={with_ln} pdl={pdl}: {e0.elapsed_time(e1) / REPS * 1000:.2f} us per iteration, {n} launches traced")
    gt = rows[:, 1:18:2].double()  # global timer per slot
    ck = rows[:, 0:18:2].double()
    for r in range(10, min(n, 13)):
        rel = (gt[r] - gt[r, 0]).tolist()
        relc = ((ck[r] - ck[r, 0]) / 1.9).tolist()
        gap = (gt[r, 0] - gt[r - 1, 8]).item()
        print(f"launch {r} sm {int(rows[r, 30])} entry-after-prev-exit {gap:7.0f} ns | " +
              " ".join(f"{nm}={relc[k]:.0f}" for k, nm in enumerate(NAMES) if k) + f" | gt exit={rel[8]:.0f}" + (" | kb1..5 " + " ".join(f"{(rows[r, 2 * k].item() - rows[r, 0].item()) / 1.9:.0f}" for k in range(9, 14)) if S else ""))
    per = (gt[11:, 8] - gt[10:-1, 8]).mean().item()
    print(f"mean exit-to-exit {per:.0f} ns")


print("step kernel slots: prologue=W issued, pdl_wait=released, tma0=first A issued, full0, mma_done=issued, acc_ready, stored=partials landed, exit")
run(1280, 1280, 64, False, pdl=1, S=4)
run(1280, 5120, 128, False, pdl=1, S=8)
