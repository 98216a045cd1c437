"""Kernel-level timing probe (CUDA events, after warm-up) through the C-ABI, every kernel next to the library that does the same
job in the reference's stack on the same box: cuBLAS (torch.matmul) for the GEMMs, cuDNN/flash SDPA for the encoder attention,
torch layer_norm, torch.stft + matmul for the log-mel.  Writes gpurun_out/perf_probe.json (committed as profiles/<tag>_library_probe.json).
Inputs are larger than L2."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from whisperjav_b200 import _lib  # noqa: E402

lib = _lib.load()
DEV = "cuda"
res = {}


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def gemm_case(M, N, K, flags=0, bias=True, res_=False, bn=0):
    A = torch.randn(M, K, device=DEV, dtype=torch.float16) * 0.5
    W = torch.randn(N, K, device=DEV, dtype=torch.float16) * 0.05
    b = torch.randn(N, device=DEV, dtype=torch.float16) if bias else None
    r = torch.randn(M, N, device=DEV, dtype=torch.float16) if res_ else None
    out = torch.empty(M, N, device=DEV, dtype=torch.float16)

    def f():
        _lib.check(lib.wjb_gemm_f16(_lib.ptr(A), K, 0, M, 1, K, _lib.ptr(W), N, K, _lib.ptr(b), _lib.ptr(r), _lib.ptr(out), N, 0, flags, bn,
                                    _lib.stream_ptr()), "gemm")
    ms = timeit(f)
    tf = 2.0 * M * N * K / ms / 1e9
    ref_ms = timeit(lambda: torch.matmul(A, W.t()))
    return {"ms": ms, "tflops": tf, "cublas_ms": ref_ms, "cublas_tflops": 2.0 * M * N * K / ref_ms / 1e9}


B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
M = B * 1500
res["gemm_qkv"] = gemm_case(M, 3840, 1280)
res["gemm_out_res"] = gemm_case(M, 1280, 1280, res_=True)
res["gemm_fc1_gelu"] = gemm_case(M, 5120, 1280, flags=1)
res["gemm_fc2_res"] = gemm_case(M, 1280, 5120, res_=True)
res["gemm_fc1_bn128"] = gemm_case(M, 5120, 1280, flags=1, bn=128)
res["gemm_dec_qkv_b64"] = gemm_case(64, 3840, 1280, bn=64)
res["gemm_dec_logits_b64"] = gemm_case(64, 51840, 1280, bias=False, bn=64)
print(json.dumps(res, indent=1), flush=True)

# attention
H, T = 20, 1500
qkv = torch.randn(B * T, 3 * 1280, device=DEV, dtype=torch.float16)
out = torch.empty(B * T, 1280, device=DEV, dtype=torch.float16)
ms = timeit(lambda: _lib.check(lib.wjb_attention_encoder_f16(_lib.ptr(qkv), _lib.ptr(out), B, T, H, _lib.stream_ptr()), "attn"))
res["attn_encoder"] = {"ms": ms, "tflops": 4.0 * B * H * T * T * 64 / ms / 1e9}
q, k, v = qkv.view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
ms2 = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
res["attn_encoder"]["torch_sdpa_ms"] = ms2

# layernorm
x = torch.randn(M, 1280, device=DEV, dtype=torch.float16)
g = torch.ones(1280, device=DEV, dtype=torch.float16)
ms = timeit(lambda: _lib.check(lib.wjb_layernorm_f16(_lib.ptr(x), _lib.ptr(g), _lib.ptr(g), _lib.ptr(out), M, 1280, _lib.stream_ptr()), "ln"))
res["layernorm"] = {"ms": ms, "GBps": 2 * M * 1280 * 2 / ms / 1e6,
                    "torch_layer_norm_ms": timeit(lambda: torch.nn.functional.layer_norm(x, (1280,), g, g))}

# cross attention decode
qd = torch.randn(B, 1280, device=DEV, dtype=torch.float16)
kv = torch.randn(B, 2 * H, T, 64, device=DEV, dtype=torch.float16)
od = torch.empty(B, 1280, device=DEV, dtype=torch.float16)
ms = timeit(lambda: _lib.check(lib.wjb_attention_cross_f16(_lib.ptr(qd), _lib.ptr(kv), _lib.ptr(od), B, H, T, _lib.stream_ptr()), "cross"), iters=20)
res["attn_cross_decode"] = {"ms": ms, "GBps": kv.numel() * 2 / ms / 1e6}

# logmel
from whisperjav_b200.model import slaney_mel_filters  # noqa: E402
audio = torch.randn(B, 480000, device=DEV) * 0.1
ns = torch.full((B,), 480000, dtype=torch.int32, device=DEV)
filt = torch.from_numpy(slaney_mel_filters(128)).to(DEV)
ws = torch.zeros(lib.wjb_logmel_workspace_bytes(B, 128), dtype=torch.uint8, device=DEV)
mel = torch.zeros(B, 3002, 128, dtype=torch.float16, device=DEV)
ms = timeit(lambda: _lib.check(lib.wjb_logmel_f16(_lib.ptr(audio), 480000, _lib.ptr(ns), B, 128, _lib.ptr(filt), _lib.ptr(mel), 1, 3002 * 128, 1, 3000, 0,
                                                  _lib.ptr(ws), _lib.stream_ptr()), "mel"))
win = torch.hann_window(400, device=DEV)
fb = filt


def torch_mel():
    st = torch.stft(audio, 400, 160, window=win, return_complex=True)
    mag = st[..., :-1].abs() ** 2
    lg = torch.clamp(fb @ mag, min=1e-10).log10()
    lg = torch.maximum(lg, lg.amax(dim=(1, 2), keepdim=True) - 8.0)
    return ((lg + 4.0) / 4.0).half()


res["logmel"] = {"ms": ms, "GBps_algorithmic": B * 2.688e6 / ms / 1e6, "torch_stft_path_ms": timeit(torch_mel)}

# VAD (Silero-class gate) over the same 64 x 30 s clips
from whisperjav_b200.vad import VadB200  # noqa: E402
vad = VadB200()
ms = timeit(lambda: vad.probs(audio, ns))
res["vad"] = {"ms": ms, "GBps_algorithmic": B * 1.92e6 / ms / 1e6, "windows": B * 938}
print(json.dumps(res, indent=1))
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/perf_probe.json").write_text(json.dumps(res, indent=1))

# scene-split energy gate over a 2 h stream (exact integer sums per 50 ms block)
from whisperjav_b200.scenes import B200SceneDetector  # noqa: E402
stream = (torch.randn(16000 * 7200, device=DEV) * 0.05)
det = B200SceneDetector()
energy = det._energy_provider(stream)
ms = timeit(lambda: energy([(0, stream.numel())], 800), iters=5)   # includes the 1.15 MB read-back of the sums
res["scene_energy"] = {"ms_incl_readback": ms, "GBps_algorithmic": stream.numel() * 4 / ms / 1e6, "blocks": stream.numel() // 800}
del stream

# decode step with every row alive: large-v3, 64 windows, EOT suppressed so that no row ever ends (steady-state ms per step)
from whisperjav_b200 import model as WM  # noqa: E402
m = WM.load_model("large-v3", max_batch=B)
xa = torch.randn(B, 1500, 1280, device=DEV, dtype=torch.float16) * 0.5
tok = WM.Tokens(m.dims.n_vocab, "ja")
for mode, kw in (("greedy", {}), ("beam2", {"beam_size": 2, "patience": 1.2})):
    m.decode_features(xa, without_timestamps=True, suppress_tokens=f"-1,{tok.eot}", sample_len=24, **kw)   # warm-up / graph capture
    torch.cuda.synchronize()
    s0 = m.stats["decode_steps"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m.decode_features(xa, without_timestamps=True, suppress_tokens=f"-1,{tok.eot}", sample_len=128, **kw)
    e1.record()
    torch.cuda.synchronize()
    steps = m.stats["decode_steps"] - s0
    res[f"decode_all_alive_{mode}"] = {"ms_per_step_incl_cross_kv_projection": e0.elapsed_time(e1) / steps, "steps": steps, "rows": B * (2 if kw else 1)}
print(json.dumps({k: res[k] for k in res if k.startswith(("scene", "decode_all"))}, indent=1))
Path("gpurun_out/perf_probe.json").write_text(json.dumps(res, indent=1))
