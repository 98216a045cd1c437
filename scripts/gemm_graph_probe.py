"""In-graph cost of the decode-step GEMM shapes: 50 dependent launches captured in a CUDA graph and replayed."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from whisperjav_b200 import _lib  # noqa: E402

lib = _lib.load()
DEV = "cuda"
M = 64
REPS = 50


def graph_time(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(REPS):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * REPS) * 1000


res = {}
for name, (N, K) in {"n1280_k1280": (1280, 1280), "qkv_3840": (3840, 1280), "fc1_5120": (5120, 1280), "fc2_k5120": (1280, 5120)}.items():
    # distinct weights per launch so the chain streams from HBM like the real 32-layer step (8 matrices cycled: > L2 for the big ones)
    Ws = [torch.randn(N, K, device=DEV, dtype=torch.float16) * 0.03 for _ in range(8)]
    A = torch.randn(M, K, device=DEV, dtype=torch.float16)
    b = torch.randn(N, device=DEV, dtype=torch.float16)
    out = torch.zeros(M, N, device=DEV, dtype=torch.float16)
    res[name] = {}
    for bn in (32, 64):
        i = [0]

        def f():
            W = Ws[i[0] % 8]
            i[0] += 1
            _lib.check(lib.wjb_gemm_f16(_lib.ptr(A), K, 0, M, 1, K, _lib.ptr(W), N, K, _lib.ptr(b), None, _lib.ptr(out), N, 0, 0, bn, _lib.stream_ptr()), "tc")
        res[name][f"tc_bn{bn}"] = round(graph_time(f), 2)
    for pdl in (0, 1):
        lib.wjb_debug_set_pdl(pdl)
        for bn, S in ((64, 2), (64, 4), (64, 8), (128, 2), (128, 4), (128, 8), (256, 2), (256, 4), (256, 8), (0, 0)):
            if bn and (N // bn * S > 148 or K // 64 < S):
                continue
            i = [0]

            def f():
                W = Ws[i[0] % 8]
                i[0] += 1
                _lib.check(lib.wjb_gemm_step_f16(_lib.ptr(A), K, M, K, _lib.ptr(W), N, K, _lib.ptr(b), None, _lib.ptr(out), N, 0, bn, S, 1,
                                                 _lib.stream_ptr()), "step")
            try:
                res[name][f"step_bn{bn}_s{S}_pdl{pdl}"] = round(graph_time(f), 2)
            except Exception as e:  # a cluster shape the device cannot co-schedule
                res[name][f"step_bn{bn}_s{S}_pdl{pdl}"] = str(e)[:60]
    lib.wjb_debug_set_pdl(0)
    i = [0]

    def g():
        W = Ws[i[0] % 8]
        i[0] += 1
        torch.matmul(A, W.t(), out=out)
    res[name]["cublas"] = round(graph_time(g), 2)
    print(name, json.dumps(res[name]), flush=True)
x = torch.randn(M, 1280, device=DEV, dtype=torch.float16)
gm = torch.ones(1280, device=DEV, dtype=torch.float16)
o = torch.empty_like(x)
res["layernorm_64x1280"] = round(graph_time(lambda: _lib.check(lib.wjb_layernorm_f16(_lib.ptr(x), _lib.ptr(gm), _lib.ptr(gm), _lib.ptr(o), M, 1280, _lib.stream_ptr()), "ln")), 2)
res["torch_layernorm"] = round(graph_time(lambda: torch.nn.functional.layer_norm(x, (1280,), gm, gm)), 2)
print(json.dumps(res))
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/gemm_graph_probe.json").write_text(json.dumps(res, indent=1))
