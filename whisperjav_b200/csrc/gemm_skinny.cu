// Weight-streaming GEMM for the decoder steps:  out[m][n] = epilogue( sum_k x[m][k] * W[n][k] ),  m <= 64.
//
// One decode step multiplies the same <= 64 activation rows (one per window in the batch) by every decoder
// weight matrix (1.6 GB per step for large-v3): HBM-bound on the weights, latency-bound on the ~200
// dependent launches.  Built for short launch-to-result time, not tensor peak:
//   * a CTA owns 8*NT output columns (NT = 4 -> 32) and one K slice; K slices of the same columns form a
//     thread-block cluster (1..8 CTAs) so that even N = 1280 fills the machine (40 column groups x 4);
//   * inside the CTA the 8 warps split the slice in 32-wide k blocks; every lane streams 16 B of one weight
//     row per column tile straight from HBM into the B fragments of mma.sync.m16n8k16 (the k index is
//     permuted identically for A and B so both come from single 16-byte loads); activations come from L2/L1
//     and are read once per (CTA, k block), i.e. N/32 times in total instead of N/8;
//   * partial sums meet in a fixed order: a 3-round smem tree over the warps, then rank 0 of the cluster
//     pulls its peers' tiles through distributed shared memory (deterministic, no atomics);
//   * bias / GELU / residual with the reference's fp16 rounding points.
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
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct SkinnyArgs {
    const __half* x;         // [M][ldx]
    const __half* W;         // [N][ldw]
    const __half* bias;      // [N] or null
    const __half* residual;  // [M][ld_out] or null (may alias out)
    __half* out;             // [M][ld_out]
    int M, N, K, ldx, ldw, ld_out, flags, ks;  // ks = cluster size = number of K slices
};

template <int MT, int NT>  // 16-row tiles (1..4), 8-column tiles per CTA (1, 2, 4)
__global__ void __launch_bounds__(kSkThreads) gemm_skinny_kernel(const SkinnyArgs a) {
    __shared__ __align__(16) float buf[4][MT * 16][NT * 8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int ks = a.ks;
    const int group = blockIdx.x / ks, rank = blockIdx.x % ks;
    const int n0 = group * (8 * NT);
    const int nkb_total = a.K / 32;
    const int kb_lo = (int)((long long)nkb_total * rank / ks), kb_hi = (int)((long long)nkb_total * (rank + 1) / ks);

    float acc[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;

    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    const __half* wrow[NT];
    bool n_ok[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int n = n0 + nt * 8 + g;
        n_ok[nt] = n < a.N;
        wrow[nt] = a.W + (size_t)(n_ok[nt] ? n : 0) * a.ldw + t4 * 8;
    }
    uint4 wv[NT], wn[NT];
    int kb = kb_lo + warp;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) wv[nt] = (kb < kb_hi && n_ok[nt]) ? __ldg(reinterpret_cast<const uint4*>(wrow[nt] + (size_t)kb * 32)) : zero4;
    while (kb < kb_hi) {
        const int kbn = kb + 8;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
            wn[nt] = (kbn < kb_hi && n_ok[nt]) ? __ldg(reinterpret_cast<const uint4*>(wrow[nt] + (size_t)kbn * 32)) : zero4;
        const __half* xk = a.x + (size_t)kb * 32 + t4 * 8;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int r0 = mt * 16 + g, r1 = r0 + 8;
            const uint4 x0 = (r0 < a.M) ? *reinterpret_cast<const uint4*>(xk + (size_t)r0 * a.ldx) : zero4;
            const uint4 x1 = (r1 < a.M) ? *reinterpret_cast<const uint4*>(xk + (size_t)r1 * a.ldx) : zero4;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                mma_16816(acc[mt][nt], x0.x, x1.x, x0.y, x1.y, wv[nt].x, wv[nt].y);
                mma_16816(acc[mt][nt], x0.z, x1.z, x0.w, x1.w, wv[nt].z, wv[nt].w);
            }
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) wv[nt] = wn[nt];
        kb = kbn;
    }

    // C fragment: c0,c1 = (row g, cols 2 t4, 2 t4 + 1), c2,c3 = (row g + 8, same cols)
    auto store_acc = [&](int slot) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                *reinterpret_cast<float2*>(&buf[slot][mt * 16 + g][nt * 8 + t4 * 2]) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
                *reinterpret_cast<float2*>(&buf[slot][mt * 16 + g + 8][nt * 8 + t4 * 2]) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
            }
    };
    auto add_acc = [&](int slot) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float2 lo = *reinterpret_cast<const float2*>(&buf[slot][mt * 16 + g][nt * 8 + t4 * 2]);
                const float2 hi = *reinterpret_cast<const float2*>(&buf[slot][mt * 16 + g + 8][nt * 8 + t4 * 2]);
                acc[mt][nt][0] += lo.x;
                acc[mt][nt][1] += lo.y;
                acc[mt][nt][2] += hi.x;
                acc[mt][nt][3] += hi.y;
            }
    };
    // warps: 8 -> 4 -> 2 -> 1, fixed order
    if (warp >= 4) store_acc(warp - 4);
    __syncthreads();
    if (warp < 4) add_acc(warp);
    __syncthreads();
    if (warp == 2 || warp == 3) store_acc(warp - 2);
    __syncthreads();
    if (warp < 2) add_acc(warp);
    __syncthreads();
    if (warp == 1) store_acc(0);
    __syncthreads();
    if (warp == 0) add_acc(0);

    if (ks > 1) {
        // K slices: every rank publishes its tile, rank 0 adds them in rank order through DSMEM
        __syncthreads();
        if (warp == 0 && rank != 0) store_acc(0);
        cluster_sync_all();
        if (warp == 0 && rank == 0) {
            for (int r = 1; r < ks; ++r) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            const uint32_t local = smem_u32(&buf[0][mt * 16 + g + 8 * hf][nt * 8 + t4 * 2]);
                            uint32_t remote;
                            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
                            float vx, vy;
                            asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(vx), "=f"(vy) : "r"(remote) : "memory");
                            acc[mt][nt][2 * hf] += vx;
                            acc[mt][nt][2 * hf + 1] += vy;
                        }
                    }
            }
        }
    }
    if (warp == 0 && rank == 0) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int m = mt * 16 + g + 8 * hf;
                    const int c0 = n0 + nt * 8 + t4 * 2;
                    if (m >= a.M || c0 >= a.N) continue;
                    const bool two = (c0 + 1 < a.N);
                    float v0 = acc[mt][nt][2 * hf], v1 = acc[mt][nt][2 * hf + 1];
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
    if (ks > 1) cluster_sync_all();  // peers keep their smem alive until rank 0 has read it
}

template <int MT, int NT>
static int launch_cfg(const SkinnyArgs& a, int groups, cudaStream_t s) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(groups * a.ks);
    cfg.blockDim = dim3(kSkThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = a.ks;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<MT, NT>, a);
    if (e != cudaSuccess) return set_error("gemm_skinny launch: %s", cudaGetErrorString(e));
    return 0;
}

static int g_force_nt = 0, g_force_ks = 0;  // tuning overrides (wjb_gemm_skinny_config); 0 = heuristic
void skinny_config(int nt, int ks) {
    g_force_nt = nt;
    g_force_ks = ks;
}

int launch_gemm_skinny(const __half* x, int ldx, const __half* W, int ldw, const __half* bias, const __half* residual, __half* out,
                       int ld_out, int M, int N, int K, int flags, cudaStream_t s) {
    if (M < 1 || M > kSkMaxM) return set_error("gemm_skinny: M=%d out of range (1..64)", M);
    if (K % 32 || ldx % 8 || ldw % 8) return set_error("gemm_skinny: K %% 32, ldx %% 8, ldw %% 8 required");
    SkinnyArgs a{x, W, bias, residual, out, M, N, K, ldx, ldw, ld_out, flags, 1};
    const int nkb = K / 32;
    // measured on B200 (scripts/skinny_sweep.py): cluster launches cost more than they save except for the long-K fc2;
    // 8 columns per CTA for <= 16 rows (x slice is small), 16 otherwise
    const int mt_ = (M + 15) / 16;
    int nt = g_force_nt ? g_force_nt : ((mt_ >= 2 && N >= 256) ? 2 : 1);
    int groups = (N + 8 * nt - 1) / (8 * nt);
    int ks = g_force_ks ? g_force_ks : ((nkb >= 128 && groups * 2 <= 3 * sm_count()) ? 2 : 1);
    a.ks = ks;
    const int mt = (M + 15) / 16;
#define WJB_SK(MT_, NT_) return launch_cfg<MT_, NT_>(a, groups, s)
#define WJB_SK_MT(NT_)           \
    switch (mt) {                \
        case 1: WJB_SK(1, NT_);  \
        case 2: WJB_SK(2, NT_);  \
        case 3: WJB_SK(3, NT_);  \
        default: WJB_SK(4, NT_); \
    }
    if (nt == 4) {
        WJB_SK_MT(4)
    } else if (nt == 2) {
        WJB_SK_MT(2)
    } else {
        WJB_SK_MT(1)
    }
#undef WJB_SK_MT
#undef WJB_SK
}

}  // namespace wjb
