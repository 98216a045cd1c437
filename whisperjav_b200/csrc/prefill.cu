// Teacher-forced decoder pass over whole token sequences ("prefill" shape): every position of every window is a GEMM row, so the
// pass costs one sweep of the decoder weights instead of one decode step per position.  Used by the word-timestamp alignment
// (openai-whisper timing.py::find_alignment runs exactly this: model.decoder(tokens, xa) with hooks on the cross-attention).
//
// Rows are laid out [B][Lp] (Lp = padded sequence length, row r = b * Lp + i); rows i >= n_tok[b] are padding: they flow through the
// GEMMs (values irrelevant) and are skipped by the attention kernels.  The Linear layers are the tcgen05 GEMM of gemm_tc.cu and the
// LayerNorms the kernel of elementwise.cu; this file holds what is new:
//   prefill_embed_kernel       token embedding + learned position, fp16-rounded fp32 sum (model.py::TextDecoder.forward)
//   prefill_self_attn_kernel   causal self-attention over <= 256 positions, one CTA per (window, head), K and V of the window in
//                              shared memory, fp32 softmax, fp16-rounded weights (the reference's rounding points)
//   prefill_cross_attn_kernel  cross-attention of 8 queries x 1500 keys per CTA: scores for the query tile stay in shared
//                              memory (fp32), softmax, fp16 weights, P.V; for the alignment layers the scaled scores are also
//                              written as fp16 (what timing.py's hooks collect)
//   prefill_prob_kernel        softmax(logits[: eot])[next token] per row (timing.py `text_token_probs`)
#include "kernels.h"

namespace wjb {

constexpr int kPfMaxL = 256;
constexpr int kPfKStride = 66;  // halfs per K/V row in shared memory (33 words: lanes reading different rows hit different banks)

__global__ void prefill_embed_kernel(const int* __restrict__ tokens, int tok_stride, const int* __restrict__ n_tok, const __half* __restrict__ emb,
                                     const __half* __restrict__ pos, __half* __restrict__ x, int Lp, int n) {
    const int r = blockIdx.x, b = r / Lp, i = r % Lp;
    __half2* o = reinterpret_cast<__half2*>(x + (long long)r * n);
    if (i >= n_tok[b]) {
        for (int k = threadIdx.x; k < n / 2; k += blockDim.x) o[k] = __floats2half2_rn(0.f, 0.f);
        return;
    }
    const int tok = tokens[(long long)b * tok_stride + i];
    const __half2* e = reinterpret_cast<const __half2*>(emb + (long long)tok * n);
    const __half2* p = reinterpret_cast<const __half2*>(pos + (long long)i * n);
    for (int k = threadIdx.x; k < n / 2; k += blockDim.x) {
        const float2 a = __half22float2(e[k]), c = __half22float2(p[k]);
        o[k] = __floats2half2_rn(a.x + c.x, a.y + c.y);
    }
}

// qkv [B*Lp][3n] (q | k | v, heads of 64) -> out [B*Lp][n]
__global__ void __launch_bounds__(256) prefill_self_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ out,
                                                                 const int* __restrict__ n_tok, int Lp, int H) {
    extern __shared__ __half pf_smem[];
    const int h = blockIdx.x, b = blockIdx.y;
    const int L = n_tok[b];
    if (L <= 0) return;
    const int n = H * 64;
    __half* Ks = pf_smem;                     // [L][kPfKStride]
    __half* Vs = Ks + (size_t)Lp * kPfKStride;  // [L][kPfKStride]
    float* sc = reinterpret_cast<float*>(Vs + (size_t)Lp * kPfKStride);  // [8 warps][Lp]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const __half* base = qkv + (long long)b * Lp * 3 * n;
    for (int k = tid; k < L * 32; k += 256) {
        const int i = k >> 5, w = k & 31;
        *reinterpret_cast<__half2*>(Ks + i * kPfKStride + 2 * w) = *reinterpret_cast<const __half2*>(base + (long long)i * 3 * n + n + h * 64 + 2 * w);
        *reinterpret_cast<__half2*>(Vs + i * kPfKStride + 2 * w) = *reinterpret_cast<const __half2*>(base + (long long)i * 3 * n + 2 * n + h * 64 + 2 * w);
    }
    __syncthreads();
    float* my = sc + warp * Lp;
    for (int i = warp; i < L; i += 8) {
        // q_i: every lane holds the whole query (64 values) in registers
        float q[64];
        const __half2* qp = reinterpret_cast<const __half2*>(base + (long long)i * 3 * n + h * 64);
#pragma unroll
        for (int d = 0; d < 32; ++d) {
            const float2 f = __half22float2(qp[d]);
            q[2 * d] = f.x;
            q[2 * d + 1] = f.y;
        }
        float mx = -INFINITY;
        for (int j = lane; j <= i; j += 32) {
            const __half2* kr = reinterpret_cast<const __half2*>(Ks + j * kPfKStride);
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < 32; ++d) {
                const float2 f = __half22float2(kr[d]);
                s = fmaf(q[2 * d], f.x, s);
                s = fmaf(q[2 * d + 1], f.y, s);
            }
            s *= 0.125f;
            my[j] = s;
            mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j <= i; j += 32) {
            const float e = __expf(my[j] - mx);
            my[j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        __syncwarp();
        // lane owns output dims 2*lane, 2*lane + 1
        float a0 = 0.f, a1 = 0.f;
        for (int j = 0; j <= i; ++j) {
            const float w = round_f16(my[j] * inv);
            const float2 v = __half22float2(*reinterpret_cast<const __half2*>(Vs + j * kPfKStride + 2 * lane));
            a0 = fmaf(w, v.x, a0);
            a1 = fmaf(w, v.y, a1);
        }
        *reinterpret_cast<__half2*>(out + ((long long)b * Lp + i) * n + h * 64 + 2 * lane) = __floats2half2_rn(a0, a1);
        __syncwarp();
    }
}

constexpr int kPfQT = 8;       // queries per CTA = warps per CTA (69 KB of shared memory: three CTAs per SM)
constexpr int kPfKC = 128;     // keys per chunk
constexpr int kPfThreads = 256;

// q [B*Lp][n], kv [B][2H][T][64] (K heads then V heads) -> out [B*Lp][n]; optional capture of the scaled scores (fp16)
__global__ void __launch_bounds__(kPfThreads) prefill_cross_attn_kernel(const __half* __restrict__ q, const __half* __restrict__ kv,
                                                                         __half* __restrict__ out, const int* __restrict__ n_tok, int Lp, int H,
                                                                         int T, const CrossCapture cap) {
    extern __shared__ __half pf_smem[];
    const int h = blockIdx.x, b = blockIdx.y, i0 = blockIdx.z * kPfQT;
    const int L = n_tok[b];
    if (i0 >= L) return;
    const int nq = min(kPfQT, L - i0);
    const int n = H * 64;
    float* sc = reinterpret_cast<float*>(pf_smem);                     // [kPfQT][T]
    __half* chunk = reinterpret_cast<__half*>(sc + (size_t)kPfQT * T);  // [kPfKC][kPfKStride]
    float* qs = reinterpret_cast<float*>(chunk + (size_t)kPfKC * kPfKStride);  // [kPfQT][64]
    float* inv = qs + kPfQT * 64;                                      // [kPfQT]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const __half* Kg = kv + ((long long)(b * 2 * H + h) * T) * 64;
    const __half* Vg = kv + ((long long)(b * 2 * H + H + h) * T) * 64;
    for (int k = tid; k < kPfQT * 64; k += kPfThreads) {
        const int qi = k >> 6, d = k & 63;
        qs[k] = qi < nq ? __half2float(q[((long long)b * Lp + i0 + qi) * n + h * 64 + d]) : 0.f;
    }
    // ---- scores: warp w handles query w; lanes take the keys of the chunk
    for (int t0 = 0; t0 < T; t0 += kPfKC) {
        const int keys = min(kPfKC, T - t0);
        __syncthreads();
        for (int k = tid; k < keys * 8; k += kPfThreads) {  // 16-byte pieces
            const int j = k >> 3, p = k & 7;
            const uint4 u = *reinterpret_cast<const uint4*>(Kg + (long long)(t0 + j) * 64 + p * 8);
            const __half2* h2 = reinterpret_cast<const __half2*>(&u);
            __half2* dst = reinterpret_cast<__half2*>(chunk + j * kPfKStride + p * 8);
#pragma unroll
            for (int e = 0; e < 4; ++e) dst[e] = h2[e];
        }
        __syncthreads();
        for (int j = lane; j < keys; j += 32) {
            float kf[64];
            const __half2* kr = reinterpret_cast<const __half2*>(chunk + j * kPfKStride);
#pragma unroll
            for (int d = 0; d < 32; ++d) {
                const float2 f = __half22float2(kr[d]);
                kf[2 * d] = f.x;
                kf[2 * d + 1] = f.y;
            }
            if (warp < nq) {
                const float* qq = qs + warp * 64;
                float s = 0.f;
#pragma unroll
                for (int d = 0; d < 64; ++d) s = fmaf(qq[d], kf[d], s);
                sc[(size_t)warp * T + t0 + j] = s * 0.125f;
            }
        }
    }
    __syncthreads();
    // ---- capture + softmax per query
    if (warp < nq) {
        const int qi = warp;
        float* row = sc + (size_t)qi * T;
        if (cap.base) {
            __half* dst = cap.base + (long long)b * cap.b_stride + (long long)h * cap.head_stride + (long long)(i0 + qi) * T;
            for (int t = lane; t < T; t += 32) dst[t] = __float2half_rn(row[t]);
        }
        float mx = -INFINITY;
        for (int t = lane; t < T; t += 32) mx = fmaxf(mx, row[t]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int t = lane; t < T; t += 32) {
            const float e = __expf(row[t] - mx);
            row[t] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        if (lane == 0) inv[qi] = 1.0f / sum;
    }
    // ---- out = P V: lane owns dims 2*lane, 2*lane + 1 of the warp's query
    float a0 = 0.f, a1 = 0.f;
    for (int t0 = 0; t0 < T; t0 += kPfKC) {
        const int keys = min(kPfKC, T - t0);
        __syncthreads();
        for (int k = tid; k < keys * 8; k += kPfThreads) {
            const int j = k >> 3, p = k & 7;
            const uint4 u = *reinterpret_cast<const uint4*>(Vg + (long long)(t0 + j) * 64 + p * 8);
            const __half2* h2 = reinterpret_cast<const __half2*>(&u);
            __half2* dst = reinterpret_cast<__half2*>(chunk + j * kPfKStride + p * 8);
#pragma unroll
            for (int e = 0; e < 4; ++e) dst[e] = h2[e];
        }
        __syncthreads();
        if (warp < nq) {
            const float* row = sc + (size_t)warp * T + t0;
            const float iv = inv[warp];
            for (int j = 0; j < keys; ++j) {
                const float w = round_f16(row[j] * iv);
                const float2 v = __half22float2(*reinterpret_cast<const __half2*>(chunk + j * kPfKStride + 2 * lane));
                a0 = fmaf(w, v.x, a0);
                a1 = fmaf(w, v.y, a1);
            }
        }
    }
    if (warp < nq) *reinterpret_cast<__half2*>(out + ((long long)b * Lp + i0 + warp) * n + h * 64 + 2 * lane) = __floats2half2_rn(a0, a1);
}

// logits fp16 [rows][stride] of rows r0 .. r0 + rows - 1 (global row index r = b * Lp + i): the row at position i predicts token i + 1
__global__ void __launch_bounds__(256) prefill_prob_kernel(const __half* __restrict__ logits, int stride, int r0, const int* __restrict__ tokens,
                                                            int tok_stride, const int* __restrict__ n_tok, float* __restrict__ prob, int Lp, int eot) {
    const int r = r0 + blockIdx.x, b = r / Lp, i = r % Lp;
    if (i + 1 >= n_tok[b]) return;
    __shared__ float red[8];
    const __half* row = logits + (long long)blockIdx.x * stride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float mx = -INFINITY;
    for (int v = tid; v < eot; v += 256) mx = fmaxf(mx, __half2float(row[v]));
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float s = 0.f;
    for (int v = tid; v < eot; v += 256) s += expf(__half2float(row[v]) - mx);
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
        float tot = 0.f;
        for (int w = 0; w < 8; ++w) tot += red[w];
        const int tok = tokens[(long long)b * tok_stride + i + 1];
        prob[(long long)b * tok_stride + i + 1] = tok < eot ? expf(__half2float(row[tok]) - mx) / tot : 0.f;
    }
}

static size_t pf_self_smem(int Lp) { return (size_t)2 * Lp * kPfKStride * 2 + (size_t)8 * Lp * 4; }
static size_t pf_cross_smem(int T) { return (size_t)kPfQT * T * 4 + (size_t)kPfKC * kPfKStride * 2 + kPfQT * 64 * 4 + kPfQT * 4 + 64; }

int prefill_init() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(prefill_self_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pf_self_smem(kPfMaxL))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(prefill_cross_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pf_cross_smem(1536))) != cudaSuccess)
        return set_error("prefill attr: %s", cudaGetErrorString(e));
    return 0;
}

int launch_prefill_embed(const int* tokens, int tok_stride, const int* n_tok, const __half* emb, const __half* pos, __half* x, int B, int Lp, int n,
                         cudaStream_t s) {
    prefill_embed_kernel<<<B * Lp, 128, 0, s>>>(tokens, tok_stride, n_tok, emb, pos, x, Lp, n);
    WJB_CHECK_LAUNCH("prefill_embed");
    return 0;
}

int launch_prefill_self_attn(const __half* qkv, __half* out, const int* n_tok, int B, int Lp, int H, cudaStream_t s) {
    if (Lp > kPfMaxL) return set_error("prefill_self_attn: %d positions > %d", Lp, kPfMaxL);
    prefill_self_attn_kernel<<<dim3(H, B), 256, pf_self_smem(Lp), s>>>(qkv, out, n_tok, Lp, H);
    WJB_CHECK_LAUNCH("prefill_self_attn");
    return 0;
}

int launch_prefill_cross_attn(const __half* q, const __half* kv, __half* out, const int* n_tok, int B, int Lp, int H, int T, const CrossCapture* cap,
                              cudaStream_t s) {
    if (T > 1536) return set_error("prefill_cross_attn: T %d > 1536", T);
    const CrossCapture c = cap ? *cap : CrossCapture{};
    prefill_cross_attn_kernel<<<dim3(H, B, (Lp + kPfQT - 1) / kPfQT), kPfThreads, pf_cross_smem(T), s>>>(q, kv, out, n_tok, Lp, H, T, c);
    WJB_CHECK_LAUNCH("prefill_cross_attn");
    return 0;
}

int launch_prefill_prob(const __half* logits, int stride, int r0, int rows, const int* tokens, int tok_stride, const int* n_tok, float* prob, int Lp,
                        int eot, cudaStream_t s) {
    if (rows <= 0) return 0;
    prefill_prob_kernel<<<rows, 256, 0, s>>>(logits, stride, r0, tokens, tok_stride, n_tok, prob, Lp, eot);
    WJB_CHECK_LAUNCH("prefill_prob");
    return 0;
}

}  // namespace wjb
