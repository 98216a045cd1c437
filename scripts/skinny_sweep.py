"""Sweep the decode-step GEMM launch configuration (columns per CTA, K slices) per shape; chain of dependent
launches timed with CUDA events.  Writes gpurun_out/skinny_sweep.json."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from whisperjav_b200 import _lib  # noqa: E402

lib = _lib.load()
DEV = "cuda"
M = 64
shapes = {"n1280_k1280": (1280, 1280), "qkv_3840": (3840, 1280), "fc1_5120": (5120, 1280), "fc2_k5120": (1280, 5120), "logits": (51866, 1280)}
res = {}
for name, (N, K) in shapes.items():
    A = torch.randn(M, K, device=DEV, dtype=torch.float16)
    W = torch.randn(N, K, device=DEV, dtype=torch.float16) * 0.03
    b = torch.randn(N, device=DEV, dtype=torch.float16)
    ld = (N + 63) // 64 * 64
    out = torch.zeros(M, ld, device=DEV, dtype=torch.float16)
    flush = torch.empty(200 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    res[name] = {}
    for nt in (1, 2, 4):
        for ks in (1, 2, 4, 8):
            if (K // 32) // ks < 4:
                continue
            lib.wjb_gemm_skinny_config(nt, ks)

            def f():
                _lib.check(lib.wjb_gemm_f16(_lib.ptr(A), K, 0, M, 1, K, _lib.ptr(W), N, K, _lib.ptr(b), None, _lib.ptr(out), ld, 0, 0, -1,
                                            _lib.stream_ptr()), "skinny")
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            reps = 30
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # weights larger than L2 are not available for a single matrix; flush L2 then time a chain
            flush.fill_(1)
            e0.record()
            for _ in range(reps):
                f()
            e1.record()
            torch.cuda.synchronize()
            res[name][f"nt{nt}_ks{ks}"] = round(e0.elapsed_time(e1) / reps * 1000, 2)  # us
    # the tcgen05 kernel with BN=64 for comparison
    def g():
        _lib.check(lib.wjb_gemm_f16(_lib.ptr(A), K, 0, M, 1, K, _lib.ptr(W), N, K, _lib.ptr(b) if N % 32 == 0 else None, None, _lib.ptr(out), ld, 0, 0, 64, _lib.stream_ptr()), "tc")
    for _ in range(3):
        g()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        g()
    e1.record()
    torch.cuda.synchronize()
    res[name]["gemm_tc_bn64"] = round(e0.elapsed_time(e1) / 30 * 1000, 2)
    res[name]["cublas"] = 0
    torch.matmul(A, W.t())
    e0.record()
    for _ in range(30):
        torch.matmul(A, W.t())
    e1.record()
    torch.cuda.synchronize()
    res[name]["cublas"] = round(e0.elapsed_time(e1) / 30 * 1000, 2)
    print(name, json.dumps(res[name]), flush=True)
lib.wjb_gemm_skinny_config(0, 0)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/skinny_sweep.json").write_text(json.dumps(res, indent=1))
