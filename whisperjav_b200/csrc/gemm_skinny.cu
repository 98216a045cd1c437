// Weight-streaming GEMM for the decoder steps:  out[m][n] = epilogue( sum_k x[m][k] * W[n][k] ),  m <= 64.
//
// One decode step multiplies the same <= 64 activation rows (one per window in the batch) by every decoder
// weight matrix (1.6 GB per step for large-v3): the work is HBM-bound on the weights and latency-bound on
// the ~200 dependent launches, so this kernel is built for short launch-to-result time, not tensor peak:
//   * one CTA per 8 output columns (N/8 CTAs: 160 .. 6484), no TMEM / TMA set-up cost;
//   * its 8 warps split K in 32-wide blocks; every lane streams 16 B of one weight row per block straight
//     from HBM into the B fragment of two mma.sync.m16n8k16 (the k index is permuted identically for A and
//     B so both come from single 16-byte loads); the activations are re-read from L2/L1;
//   * partial sums meet in smem in a fixed order (deterministic), then bias / GELU / residual with the
//     reference's fp16 rounding points.
// Replaces the per-token Linear calls of openai-whisper model.py::TextDecoder (cuBLAS GEMV/GEMM there).
#include "kernels.h"

namespace wjb {

constexpr int kSkThreads = 256;
constexpr int kSkMaxM = 64;

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct SkinnyArgs {
    const __half* x;      // [M][ldx]
    const __half* W;      // [N][ldw]
    const __half* bias;   // [N] or null
    const __half* residual;  // [M][ld_out] or null (may alias out)
    __half* out;          // [M][ld_out]
    int M, N, K, ldx, ldw, ld_out, flags;
};

template <int MT>  // number of 16-row tiles (1..4)
__global__ void __launch_bounds__(kSkThreads) gemm_skinny_kernel(const SkinnyArgs a) {
    __shared__ float part[8][MT * 16][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int n0 = blockIdx.x * 8;
    const int n = n0 + g;                       // weight row this lane streams
    const bool n_ok = n < a.N;
    const int nkb = a.K / 32;
    float acc[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][i] = 0.f;

    const __half* wrow = a.W + (size_t)(n_ok ? n : 0) * a.ldw + t4 * 8;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    // software prefetch of the weight stream: three k-blocks in flight per warp
    auto wload = [&](int k) { return (k < nkb && n_ok) ? __ldg(reinterpret_cast<const uint4*>(wrow + (size_t)k * 32)) : zero4; };
    int kb = warp;
    uint4 wv = wload(kb), w1 = wload(kb + 8);
    while (kb < nkb) {
        const int kb_next = kb + 8;
        const uint4 wnext = w1;
        w1 = wload(kb + 16);
        const __half* xk = a.x + (size_t)kb * 32 + t4 * 8;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int r0 = mt * 16 + g, r1 = r0 + 8;
            const uint4 x0 = (r0 < a.M) ? *reinterpret_cast<const uint4*>(xk + (size_t)r0 * a.ldx) : zero4;
            const uint4 x1 = (r1 < a.M) ? *reinterpret_cast<const uint4*>(xk + (size_t)r1 * a.ldx) : zero4;
            mma_16816(acc[mt], x0.x, x1.x, x0.y, x1.y, wv.x, wv.y);
            mma_16816(acc[mt], x0.z, x1.z, x0.w, x1.w, wv.z, wv.w);
        }
        wv = wnext;
        kb = kb_next;
    }
    // partial tiles -> smem; C fragment: c0,c1 = (row g, cols 2 t4, 2 t4 + 1), c2,c3 = (row g + 8, same cols)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        part[warp][mt * 16 + g][t4 * 2] = acc[mt][0];
        part[warp][mt * 16 + g][t4 * 2 + 1] = acc[mt][1];
        part[warp][mt * 16 + g + 8][t4 * 2] = acc[mt][2];
        part[warp][mt * 16 + g + 8][t4 * 2 + 1] = acc[mt][3];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < MT * 16 * 4; i += kSkThreads) {
        const int m = i >> 2, cp = (i & 3) * 2;  // two adjacent columns per thread
        if (m >= a.M) continue;
        float v0 = 0.f, v1 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            v0 += part[w][m][cp];
            v1 += part[w][m][cp + 1];
        }
        const int c0 = n0 + cp;
        if (c0 >= a.N) continue;
        const bool two = (c0 + 1 < a.N);
        if (a.bias) {
            v0 += __half2float(a.bias[c0]);
            if (two) v1 += __half2float(a.bias[c0 + 1]);
        }
        v0 = round_f16(v0);
        v1 = round_f16(v1);
        if (a.flags & GEMM_GELU) {
            v0 = round_f16(gelu_erf(v0));
            v1 = round_f16(gelu_erf(v1));
        }
        const size_t off = (size_t)m * a.ld_out + c0;
        if (a.residual) {
            v0 += __half2float(a.residual[off]);
            if (two) v1 += __half2float(a.residual[off + 1]);
        }
        if (two && ((off & 1) == 0)) {
            *reinterpret_cast<__half2*>(a.out + off) = __floats2half2_rn(v0, v1);
        } else {
            a.out[off] = __float2half_rn(v0);
            if (two) a.out[off + 1] = __float2half_rn(v1);
        }
    }
}

int launch_gemm_skinny(const __half* x, int ldx, const __half* W, int ldw, const __half* bias, const __half* residual, __half* out,
                       int ld_out, int M, int N, int K, int flags, cudaStream_t s) {
    if (M < 1 || M > kSkMaxM) return set_error("gemm_skinny: M=%d out of range (1..64)", M);
    if (K % 32 || ldx % 8 || ldw % 8) return set_error("gemm_skinny: K %% 32, ldx %% 8, ldw %% 8 required");
    SkinnyArgs a{x, W, bias, residual, out, M, N, K, ldx, ldw, ld_out, flags};
    const int grid = (N + 7) / 8;
    const int mt = (M + 15) / 16;
    switch (mt) {
        case 1: gemm_skinny_kernel<1><<<grid, kSkThreads, 0, s>>>(a); break;
        case 2: gemm_skinny_kernel<2><<<grid, kSkThreads, 0, s>>>(a); break;
        case 3: gemm_skinny_kernel<3><<<grid, kSkThreads, 0, s>>>(a); break;
        default: gemm_skinny_kernel<4><<<grid, kSkThreads, 0, s>>>(a); break;
    }
    WJB_CHECK_LAUNCH("gemm_skinny");
    return 0;
}

}  // namespace wjb
