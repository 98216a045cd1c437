"""Kernel-level parity on the GPU, through the C-ABI (ctypes), against plain torch fp32 math on the
same fp16 inputs.  Tolerances are written next to each assertion."""
import ctypes as C
import json

import numpy as np
import pytest
import torch

from whisperjav_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


def r16(t):
    return t.half().float()


def gemm(lib, A, W, bias=None, res=None, flags=0, block_n=0, rows_per_batch=None, n_batch=1, a_row_stride=None,
         a_batch_stride=0, K=None, out=None):
    N = W.shape[0]
    K = K or W.shape[1]
    rows = rows_per_batch or A.shape[0]
    if out is None:
        out = torch.empty(n_batch * rows, N, dtype=torch.float16, device=DEV)
    _lib.check(lib.wjb_gemm_f16(_lib.ptr(A), a_row_stride or K, a_batch_stride, rows, n_batch, K, _lib.ptr(W), N, W.shape[1],
                                _lib.ptr(bias), _lib.ptr(res), _lib.ptr(out), N, rows * N, flags, block_n, _lib.stream_ptr()), "gemm")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("M,N,K,bn", [(128, 256, 64, 256), (128, 256, 256, 256), (256, 512, 128, 128), (1000, 1280, 1280, 0),
                                      (64, 1280, 1280, 64), (64, 1280, 1280, 32), (16, 3840, 1280, 32), (64, 1280, 5120, 0), (3, 384, 384, 64), (130, 51866, 384, 0), (4096, 3840, 1280, 256)])
def test_gemm_plain(lib, diag_dir, M, N, K, bn):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    A = (torch.randn(M, K, generator=g) * 0.5).half().to(DEV)
    W = (torch.randn(N, K, generator=g) * 0.05).half().to(DEV)
    if N % 64:
        outbuf = torch.zeros(M, (N + 63) // 64 * 64, dtype=torch.float16, device=DEV)
        _lib.check(lib.wjb_gemm_f16(_lib.ptr(A), K, 0, M, 1, K, _lib.ptr(W), N, K, None, None, _lib.ptr(outbuf), outbuf.shape[1], 0, 0, bn,
                                    _lib.stream_ptr()), "gemm")
        torch.cuda.synchronize()
        out = outbuf[:, :N]
    else:
        out = gemm(lib, A, W, block_n=bn)
    ref = A.float() @ W.float().t()
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    if not err <= 2e-3 * scale + 1e-3:
        np.save(diag_dir / f"gemm_{M}_{N}_{K}_out.npy", out[:256, :512].float().cpu().numpy())
        np.save(diag_dir / f"gemm_{M}_{N}_{K}_ref.npy", ref[:256, :512].cpu().numpy())
    # fp16 output rounding: <= 2^-11 relative of the value plus accumulation-order noise
    assert err <= 2e-3 * scale + 1e-3, (err, scale)


def test_gemm_epilogues(lib):
    g = torch.Generator().manual_seed(3)
    M, N, K = 777, 1280, 640
    A = (torch.randn(M, K, generator=g)).half().to(DEV)
    W = (torch.randn(N, K, generator=g) * 0.04).half().to(DEV)
    bias = (torch.randn(N, generator=g) * 0.3).half().to(DEV)
    res = (torch.randn(M, N, generator=g)).half().to(DEV)
    lin = r16(A.float() @ W.float().t() + bias.float())
    out = gemm(lib, A, W, bias=bias)
    assert (out.float() - lin).abs().max().item() <= 4e-3
    out = gemm(lib, A, W, bias=bias, flags=1)
    ref = r16(torch.nn.functional.gelu(lin))
    assert (out.float() - ref).abs().max().item() <= 4e-3
    out = gemm(lib, A, W, bias=bias, res=res)
    ref = r16(lin + res.float())
    assert (out.float() - ref).abs().max().item() <= 8e-3
    # in-place residual (out aliases residual), as the transformer blocks use it
    x = res.clone()
    gemm(lib, A, W, bias=bias, res=x, out=x)
    assert (x.float() - ref).abs().max().item() <= 8e-3

@pytest.mark.parametrize("M,N,K,bn,splits", [(64, 1280, 1280, 64, -1), (64, 1280, 1280, 32, 3), (64, 3840, 1280, 64, 2), (33, 5120, 1280, 64, -1),
                                             (64, 1280, 5120, 64, 7), (1, 1280, 1280, 64, -1), (128, 1280, 5120, 128, 5), (100, 384, 384, 64, 3)])
def test_gemm_splitk_decode(lib, M, N, K, bn, splits):
    """tcgen05 GEMM with K sliced over CTAs (decoder steps): every epilogue, repeated launches on one workspace, and
    bit-identical results from run to run (the slices are added in a fixed order by the last CTA to arrive)."""
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g)).half().to(DEV)
    W = (torch.randn(N, K, generator=g) * 0.03).half().to(DEV)
    bias = (torch.randn(N, generator=g) * 0.3).half().to(DEV)
    res = (torch.randn(M, N, generator=g)).half().to(DEV)
    ws_bytes = lib.wjb_gemm_splitk_workspace_bytes()
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=DEV)
    lin = r16(A.float() @ W.float().t() + bias.float())
    for flags, use_res in ((0, False), (1, False), (0, True)):
        outs = []
        for rep in range(3):
            out = torch.zeros(M, N, dtype=torch.float16, device=DEV)
            rbuf = None
            if use_res:
                out.copy_(res)
                rbuf = out  # in place, as the decoder uses it
            _lib.check(lib.wjb_gemm_f16_splitk(_lib.ptr(A), K, M, K, _lib.ptr(W), N, K, _lib.ptr(bias), _lib.ptr(rbuf), _lib.ptr(out), N, flags, bn,
                                               splits, _lib.ptr(ws), ws_bytes, _lib.stream_ptr()), "split-K gemm")
            torch.cuda.synchronize()
            outs.append(out)
        ref = lin
        if flags:
            ref = r16(torch.nn.functional.gelu(lin))
        if use_res:
            ref = r16(lin + res.float())
        err = (outs[0].float() - ref).abs().max().item()
        assert err <= 8e-3 * max(1.0, ref.abs().max().item() / 4), (flags, use_res, err)
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert int(ws[:4096].view(torch.int32).abs().sum().item()) == 0  # the counters are back at zero

@pytest.mark.parametrize("M,N,K,bn,S", [(64, 1280, 1280, 0, 0), (64, 1280, 1280, 64, 4), (64, 1280, 1280, 128, 8), (64, 3840, 1280, 256, 8),
                                        (33, 5120, 1280, 256, 4), (64, 1280, 5120, 128, 8), (1, 1280, 1280, 64, 2), (128, 1280, 2560, 0, 0),
                                        (100, 384, 768, 64, 2), (128, 3840, 1280, 128, 4), (64, 5120, 1280, 0, 0), (7, 384, 512, 0, 0)])
def test_gemm_step_cluster(lib, M, N, K, bn, S):
    """Decoder-step GEMM: cluster split-K through distributed shared memory, every epilogue, in place residual, repeatable bits."""
    g = torch.Generator().manual_seed(M + N + K + bn + S)
    A = (torch.randn(M, K, generator=g)).half().to(DEV)
    W = (torch.randn(N, K, generator=g) * 0.03).half().to(DEV)
    bias = (torch.randn(N, generator=g) * 0.3).half().to(DEV)
    res = (torch.randn(M, N, generator=g)).half().to(DEV)
    lin = r16(A.float() @ W.float().t() + bias.float())
    for flags, use_res, early in ((0, False, 0), (1, False, 1), (0, True, 1)):
        outs = []
        for rep in range(2):
            out = torch.zeros(M, N, dtype=torch.float16, device=DEV)
            rbuf = None
            if use_res:
                out.copy_(res)
                rbuf = out
            _lib.check(lib.wjb_gemm_step_f16(_lib.ptr(A), K, M, K, _lib.ptr(W), N, K, _lib.ptr(bias), _lib.ptr(rbuf), _lib.ptr(out), N, flags, bn, S,
                                             early, _lib.stream_ptr()), "step gemm")
            torch.cuda.synchronize()
            outs.append(out)
        ref = lin
        if flags:
            ref = r16(torch.nn.functional.gelu(lin))
        if use_res:
            ref = r16(lin + res.float())
        err = (outs[0].float() - ref).abs().max().item()
        assert err <= 8e-3 * max(1.0, ref.abs().max().item() / 4), (flags, use_res, err)
        assert torch.equal(outs[0], outs[1])

@pytest.mark.parametrize("M,N,K,bn,S", [(64, 3840, 1280, 0, 0), (64, 1280, 1280, 64, 4), (64, 5120, 1280, 0, 0), (33, 384, 384, 0, 0),
                                        (128, 1280, 1280, 0, 0), (100, 1152, 384, 64, 2), (1, 1280, 1280, 128, 8), (64, 1280, 1280, 64, 1)])
def test_gemm_step_fused_layernorm(lib, M, N, K, bn, S):
    """LayerNorm + Linear in one launch (per-row statistics exchanged across the cluster) against LayerNorm -> fp16 -> Linear."""
    g = torch.Generator().manual_seed(M + N + K + bn + S)
    A = (torch.randn(M, K, generator=g) * 1.7 + 0.3).half().to(DEV)
    W = (torch.randn(N, K, generator=g) * 0.03).half().to(DEV)
    bias = (torch.randn(N, generator=g) * 0.3).half().to(DEV)
    gamma = (1.0 + 0.2 * torch.randn(K, generator=g)).half().to(DEV)
    beta = (0.1 * torch.randn(K, generator=g)).half().to(DEV)
    res = (torch.randn(M, N, generator=g)).half().to(DEV)
    h = r16(torch.nn.functional.layer_norm(A.float(), (K,), gamma.float(), beta.float(), 1e-5))
    lin = r16(h @ W.float().t() + bias.float())
    a_before = A.clone()
    for flags, use_res in ((0, False), (1, False), (0, True)):
        outs = []
        for rep in range(2):
            out = torch.zeros(M, N, dtype=torch.float16, device=DEV)
            rbuf = None
            if use_res:
                out.copy_(res)
                rbuf = out
            _lib.check(lib.wjb_gemm_step_ln_f16(_lib.ptr(A), K, M, K, _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(W), N, K, _lib.ptr(bias), _lib.ptr(rbuf),
                                                _lib.ptr(out), N, flags, bn, S, 1, _lib.stream_ptr()), "step gemm + ln")
            torch.cuda.synchronize()
            outs.append(out)
        ref = lin
        if flags:
            ref = r16(torch.nn.functional.gelu(lin))
        if use_res:
            ref = r16(lin + res.float())
        err = (outs[0].float() - ref).abs().max().item()
        # one fp16 ulp of a normalised activation (2^-10 at |h| ~ 2) times |W| summed over K stays well inside this
        assert err <= 1.2e-2 * max(1.0, ref.abs().max().item() / 4), (flags, use_res, err)
        assert torch.equal(outs[0], outs[1])
    assert torch.equal(A, a_before)  # the residual stream itself is not touched


@pytest.mark.parametrize("M,N,K,bn,S", [(64, 1280, 1280, 0, 0), (64, 1280, 5120, 0, 0), (64, 1280, 1280, 64, 4), (128, 1280, 1280, 128, 8),
                                        (96, 1280, 1280, 0, 0), (33, 384, 1536, 0, 0), (7, 384, 384, 64, 2)])
def test_gemm_step_row_statistics_chain(lib, M, N, K, bn, S):
    """The decode step's LayerNorm without a launch: a residual-producing step GEMM adds fixed-point row statistics of what it stores,
    the next step GEMM normalises its activation tile from them.  (i) the statistics are the exact fixed-point sums of the stored
    fp16 rows up to the rounding of the per-CTA fp32 partials, identical run to run; (ii) LayerNorm + Linear through them equals
    LayerNorm -> fp16 -> Linear (torch fp32) like the exchanging kernel; (iii) a wrong statistic changes the result (it is used)."""
    g = torch.Generator().manual_seed(M * 7 + N + K + bn + S)
    A = (torch.randn(M, K, generator=g) * 0.7).half().to(DEV)
    W = (torch.randn(N, K, generator=g) * (1.0 / K ** 0.5)).half().to(DEV)
    bias = (torch.randn(N, generator=g) * 0.3).half().to(DEV)
    res = (torch.randn(M, N, generator=g) * 2.0 + 0.4).half().to(DEV)
    xs, sts = [], []
    for rep in range(2):
        x = res.clone()
        st = torch.zeros(M, 2, dtype=torch.int64, device=DEV)
        _lib.check(lib.wjb_gemm_step_stats_f16(_lib.ptr(A), K, M, K, None, None, None, _lib.ptr(W), N, K, _lib.ptr(bias), _lib.ptr(x), _lib.ptr(x), N,
                                               _lib.ptr(st), 0, bn, S, 1, _lib.stream_ptr()), "producer")
        torch.cuda.synchronize()
        xs.append(x)
        sts.append(st)
    assert torch.equal(xs[0], xs[1]) and torch.equal(sts[0], sts[1])
    ref_x = r16(r16(A.float() @ W.float().t() + bias.float()) + res.float())
    assert (xs[0].float() - ref_x).abs().max().item() <= 8e-3 * max(1.0, ref_x.abs().max().item() / 4)
    xd = xs[0].double()
    want = torch.stack([xd.sum(1), (xd * xd).sum(1)], 1) * 2 ** 20
    # sums of fp16 values: exact integers at the 2^20 scale, except that the few elements below 2^-10 round to the grid (<= 0.5 each)
    assert (sts[0][:, 0].double() - want[:, 0]).abs().max().item() <= 4.0
    err = (sts[0][:, 1].double() - want[:, 1]).abs().max().item()
    assert err <= 0.5 * N, err                                          # every square rounded to the 2^-20 grid once, then exact adds
    # consumer: N2 columns out of the normalised stream
    K2, N2 = N, 3 * N if N <= 1280 else N
    W2 = (torch.randn(N2, K2, generator=g) * 0.03).half().to(DEV)
    b2 = (torch.randn(N2, generator=g) * 0.3).half().to(DEV)
    gamma = (1.0 + 0.2 * torch.randn(K2, generator=g)).half().to(DEV)
    beta = (0.1 * torch.randn(K2, generator=g)).half().to(DEV)
    h = r16(torch.nn.functional.layer_norm(xs[0].float(), (K2,), gamma.float(), beta.float(), 1e-5))
    ref = r16(h @ W2.float().t() + b2.float())
    outs = []
    for st in (sts[0], sts[0], sts[0] + torch.tensor([[1 << 27, 0]], device=DEV)):
        out = torch.zeros(M, N2, dtype=torch.float16, device=DEV)
        _lib.check(lib.wjb_gemm_step_stats_f16(_lib.ptr(xs[0]), K2, M, K2, _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(st), _lib.ptr(W2), N2, K2, _lib.ptr(b2),
                                               None, _lib.ptr(out), N2, None, 0, 0, 0, 1, _lib.stream_ptr()), "consumer")
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    assert (outs[0].float() - ref).abs().max().item() <= 1.2e-2 * max(1.0, ref.abs().max().item() / 4)
    assert not torch.equal(outs[0], outs[2])
    # and it agrees with the stand-alone LayerNorm kernel + plain step GEMM to the last bit almost everywhere (same rounding points)
    hk = torch.empty_like(xs[0])
    _lib.check(lib.wjb_layernorm_f16(_lib.ptr(xs[0]), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(hk), M, K2, _lib.stream_ptr()), "ln")
    out2 = torch.zeros(M, N2, dtype=torch.float16, device=DEV)
    _lib.check(lib.wjb_gemm_step_f16(_lib.ptr(hk), K2, M, K2, _lib.ptr(W2), N2, K2, _lib.ptr(b2), None, _lib.ptr(out2), N2, 0, 0, 0, 1, _lib.stream_ptr()), "plain")
    torch.cuda.synchronize()
    assert (outs[0].float() - out2.float()).abs().max().item() <= 4e-3 * max(1.0, ref.abs().max().item() / 4)
    assert (outs[0] == out2).float().mean().item() >= 0.98


@pytest.mark.parametrize("C_,stride,T_out", [(128, 1, 3000), (80, 1, 3000), (384, 2, 1500), (1280, 2, 1500)])
def test_gemm_conv_im2col_free(lib, C_, stride, T_out):
    """k=3 convolution as a GEMM over overlapping rows of the channels-last padded input."""
    B, N = 2, 256
    T_in = T_out * stride
    g = torch.Generator().manual_seed(C_)
    x = (torch.randn(B, C_, T_in, generator=g)).half()
    w = (torch.randn(N, C_, 3, generator=g) * (1.0 / (3 * C_) ** 0.5)).half()
    bias = (torch.randn(N, generator=g) * 0.1).half()
    xpad = torch.zeros(B, T_in + 2, C_, dtype=torch.float16)
    xpad[:, 1:-1] = x.permute(0, 2, 1)
    wk = w.permute(0, 2, 1).reshape(N, 3 * C_).contiguous()
    out = torch.empty(B * T_out, N, dtype=torch.float16, device=DEV)
    xp, wkd, bd = xpad.to(DEV), wk.to(DEV), bias.to(DEV)
    _lib.check(lib.wjb_gemm_f16(_lib.ptr(xp), stride * C_, (T_in + 2) * C_, T_out, B, 3 * C_, _lib.ptr(wkd), N, 3 * C_, _lib.ptr(bd), None,
                                _lib.ptr(out), N, T_out * N, 0, 0, _lib.stream_ptr()), "conv gemm")
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv1d(x.float(), w.float(), bias.float(), stride=stride, padding=1).permute(0, 2, 1).reshape(B * T_out, N)
    err = (out.float().cpu() - ref).abs().max().item()
    assert err <= 4e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("rows,n", [(5, 384), (1000, 1280), (64, 1280), (33, 512)])
def test_layernorm(lib, rows, n):
    g = torch.Generator().manual_seed(rows)
    x = (torch.randn(rows, n, generator=g) * 3 + 0.5).half().to(DEV)
    ga = (1 + 0.1 * torch.randn(n, generator=g)).half().to(DEV)
    be = (0.1 * torch.randn(n, generator=g)).half().to(DEV)
    out = torch.empty_like(x)
    _lib.check(lib.wjb_layernorm_f16(_lib.ptr(x), _lib.ptr(ga), _lib.ptr(be), _lib.ptr(out), rows, n, _lib.stream_ptr()), "ln")
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.float(), (n,), ga.float(), be.float(), 1e-5)
    assert (out.float() - ref).abs().max().item() <= 4e-3  # one fp16 ulp at |y| < 4


@pytest.mark.parametrize("B,T,H,scale", [(1, 1500, 6, 1.2), (2, 1500, 20, 1.2), (1, 256, 2, 1.2), (3, 200, 1, 1.2),
                                           (1, 1500, 2, 3.0), (1, 1500, 2, 0.05)])
def test_attention_encoder(lib, diag_dir, B, T, H, scale):
    """scale 3.0 makes score maxima jump by more than 2^8 between key blocks (exercises the lazy O rescale);
    scale 0.05 gives a nearly uniform softmax."""
    n = 64 * H
    g = torch.Generator().manual_seed(T + H)
    qkv = (torch.randn(B * T, 3 * n, generator=g) * scale).half().to(DEV)
    out = torch.zeros(B * T, n, dtype=torch.float16, device=DEV)
    _lib.check(lib.wjb_attention_encoder_f16(_lib.ptr(qkv), _lib.ptr(out), B, T, H, _lib.stream_ptr()), "attn")
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v
    ref = ref.permute(0, 2, 1, 3).reshape(B * T, n)
    err = (out.float() - ref).abs().max().item()
    tol = 6e-3 * max(1.0, scale)
    if not err <= tol:
        np.save(diag_dir / f"attn_{B}_{T}_{H}_{scale}_out.npy", out[:256].float().cpu().numpy())
        np.save(diag_dir / f"attn_{B}_{T}_{H}_{scale}_ref.npy", ref[:256].cpu().numpy())
    # fp16 P and fp16 output: a few 1e-3 absolute on O(1) values
    assert err <= tol, err

@pytest.mark.parametrize("B,H,pos", [(2, 6, 0), (3, 6, 5), (64, 20, 226), (2, 1, 447), (5, 20, 63), (4, 6, 64)])
def test_attention_self_decode(lib, B, H, pos):
    """Decoder self-attention at one position: cache append + single-query attention over 0..pos."""
    n, n_ctx = 64 * H, 448
    g = torch.Generator().manual_seed(B + H + pos)
    qkv = (torch.randn(B, 3 * n, generator=g) * 1.2).half().to(DEV)
    cache = (torch.randn(B, 2 * H, n_ctx, 64, generator=g)).half().to(DEV)
    before = cache.clone()
    out = torch.empty(B, n, dtype=torch.float16, device=DEV)
    p = torch.tensor([pos], dtype=torch.int32, device=DEV)
    _lib.check(lib.wjb_attention_self_f16(_lib.ptr(qkv), _lib.ptr(cache), _lib.ptr(out), _lib.ptr(p), B, H, n_ctx, _lib.stream_ptr()), "self")
    torch.cuda.synchronize()
    k_new = qkv[:, n:2 * n].view(B, H, 64)
    v_new = qkv[:, 2 * n:].view(B, H, 64)
    assert torch.equal(cache[:, :H, pos], k_new) and torch.equal(cache[:, H:, pos], v_new)
    keep = torch.ones(n_ctx, dtype=torch.bool)
    keep[pos] = False
    assert torch.equal(cache[:, :, keep], before[:, :, keep])  # nothing else is touched
    q = qkv[:, :n].float().view(B, H, 1, 64)
    k, v = cache[:, :H, : pos + 1].float(), cache[:, H:, : pos + 1].float()
    w = r16(torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1))
    ref = (w @ v).reshape(B, n)
    assert (out.float() - ref).abs().max().item() <= 3e-3


@pytest.mark.parametrize("B,H,T", [(2, 6, 1500), (64, 20, 1500), (3, 1, 100)])
def test_attention_cross(lib, B, H, T):
    n = 64 * H
    g = torch.Generator().manual_seed(B + T)
    q = (torch.randn(B, n, generator=g) * 1.5).half().to(DEV)
    kv = (torch.randn(B, 2 * H, T, 64, generator=g)).half().to(DEV)
    out = torch.empty(B, n, dtype=torch.float16, device=DEV)
    _lib.check(lib.wjb_attention_cross_f16(_lib.ptr(q), _lib.ptr(kv), _lib.ptr(out), B, H, T, _lib.stream_ptr()), "cross")
    torch.cuda.synchronize()
    qh = q.float().view(B, H, 1, 64)
    k, v = kv[:, :H].float(), kv[:, H:].float()
    w = r16(torch.softmax(qh @ k.transpose(-1, -2) * 0.125, -1))
    ref = (w @ v).reshape(B, n)
    assert (out.float() - ref).abs().max().item() <= 3e-3


@pytest.mark.parametrize("Wn,beams,H,T", [(3, 2, 6, 1500), (64, 2, 20, 1500), (5, 3, 20, 1500), (4, 4, 6, 700), (3, 5, 6, 1500)])
def test_attention_cross_beams_share_the_window_stream(lib, Wn, beams, H, T):
    """Beam search: the queries of a window's beams against that window's K/V (one 256-thread CTA streams them once for beams <= 4,
    beams = 5 falls back to one CTA per row); same math and tolerance as the single-query kernel, identical run to run."""
    n = 64 * H
    g = torch.Generator().manual_seed(Wn * 10 + beams)
    q = (torch.randn(Wn * beams, n, generator=g) * 1.5).half().to(DEV)
    kv = (torch.randn(Wn, 2 * H, T, 64, generator=g)).half().to(DEV)
    outs = []
    for _ in range(2):
        out = torch.zeros(Wn * beams, n, dtype=torch.float16, device=DEV)
        _lib.check(lib.wjb_attention_cross_beam_f16(_lib.ptr(q), _lib.ptr(kv), _lib.ptr(out), Wn * beams, H, T, beams, _lib.stream_ptr()), "cross beam")
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    qh = q.float().view(Wn, beams, H, 64).permute(0, 2, 1, 3)                 # [window][head][beam][64]
    k, v = kv[:, :H].float(), kv[:, H:].float()
    w = r16(torch.softmax(qh @ k.transpose(-1, -2) * 0.125, -1))
    ref = (w @ v).permute(0, 2, 1, 3).reshape(Wn * beams, n)
    assert (outs[0].float() - ref).abs().max().item() <= 3e-3
    # a window's beams fed one at a time through the single-query kernel: the same numbers up to the fp32 summation order
    single = torch.zeros(Wn, n, dtype=torch.float16, device=DEV)
    q0 = q.view(Wn, beams, n)[:, 0].contiguous()
    _lib.check(lib.wjb_attention_cross_f16(_lib.ptr(q0), _lib.ptr(kv), _lib.ptr(single), Wn, H, T, _lib.stream_ptr()), "cross")
    torch.cuda.synchronize()
    assert (outs[0].view(Wn, beams, n)[:, 0].float() - single.float()).abs().max().item() <= 2e-3


@pytest.mark.parametrize("n_mels", [80, 128])
def test_logmel_vs_torch(lib, diag_dir, n_mels):
    """Fused STFT+mel+log kernel vs torch.stft in fp32 (tolerance from north_star: 1e-3 abs)."""
    from whisperjav_b200.model import slaney_mel_filters
    from whisperjav_b200.synth import speech_shaped_audio
    clips = [speech_shaped_audio(s, 50 + i) for i, s in enumerate([30.0, 7.3, 0.5, 29.99])]
    clips.append(np.zeros(16000, np.float32))
    B = len(clips)
    S = max(len(c) for c in clips)
    audio = torch.zeros(B, S)
    for i, c in enumerate(clips):
        audio[i, : len(c)] = torch.from_numpy(c)
    ns = torch.tensor([len(c) for c in clips], dtype=torch.int32)
    filt = torch.from_numpy(slaney_mel_filters(n_mels))
    for layout in ("mel", "time"):
        nf = 3000
        a_d, ns_d, f_d = audio.to(DEV), ns.to(DEV), filt.to(DEV)
        ws = torch.zeros(lib.wjb_logmel_workspace_bytes(B, n_mels), dtype=torch.uint8, device=DEV)
        if layout == "mel":
            out = torch.full((B, n_mels, nf), 7.0, dtype=torch.float16, device=DEV)
            args = (0, n_mels * nf, 0)
        else:
            out = torch.zeros(B, nf + 2, n_mels, dtype=torch.float16, device=DEV)
            args = (1, (nf + 2) * n_mels, 1)
        _lib.check(lib.wjb_logmel_f16(_lib.ptr(a_d), S, _lib.ptr(ns_d), B, n_mels, _lib.ptr(f_d), _lib.ptr(out), args[0], args[1], args[2], nf, 0,
                                      _lib.ptr(ws), _lib.stream_ptr()), "logmel")
        torch.cuda.synchronize()
        got = out.float().cpu()
        if layout == "time":
            assert got[:, 0].abs().max() == 0 and got[:, -1].abs().max() == 0
            got = got[:, 1:-1].permute(0, 2, 1)
        for i, c in enumerate(clips):
            x = torch.nn.functional.pad(torch.from_numpy(c), (0, 480000))
            st = torch.stft(x, 400, 160, window=torch.hann_window(400), return_complex=True)
            mag = st[..., :-1].abs() ** 2
            ls = torch.clamp(filt @ mag, min=1e-10).log10()
            ls = (torch.maximum(ls, ls.max() - 8.0) + 4.0) / 4.0
            cf = len(c) // 160
            ref = torch.zeros(n_mels, nf)
            ref[:, : min(cf, nf)] = ls[:, : min(cf, nf)]
            err = (got[i] - ref).abs().max().item()
            if err > 1e-3:
                np.save(diag_dir / f"mel_{n_mels}_{layout}_{i}_got.npy", got[i].numpy())
                np.save(diag_dir / f"mel_{n_mels}_{layout}_{i}_ref.npy", ref.numpy())
            assert err <= 1e-3, (layout, i, err)
