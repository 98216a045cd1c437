// Persistent decoder-step kernel: ONE launch runs a whole decoder step for <= 64 windows.
//
// The per-step work of openai-whisper's TextDecoder (embedding, 32 x [LN, QKV, cached self-attention, out-proj,
// LN, cross-attention over 1500 frames, out-proj, LN, MLP], final LN, tied-embedding logits) is ~355 dependent
// kernels when launched one by one, each a few microseconds of latency-bound work: launch and drain gaps, not HBM
// bandwidth, set the step time.  Here one CTA per SM stays resident and the phases are separated by a grid-wide
// barrier (one atomic arrival + acquire spin per CTA), so the 17.3 GB a step has to move (1.6 GB of weights once,
// 245.76 MB of cross-K/V per window) stream with the whole machine participating in every phase.
//
//   phase kinds   GEMM  : out[64][N] = epi(A[64][K] W[N][K]^T); 8-column tiles dealt round-robin to CTAs, a CTA
//                         works on up to 4 tiles at once, its 8 warps split K in 32-wide blocks (4 weight loads of
//                         16 B per lane in flight), partial sums meet in a fixed smem tree, warp 0 applies
//                         bias / GELU / residual with the reference's fp16 rounding points
//                 LN    : one warp per row (fp32 two-pass statistics), 64 rows dealt to CTAs
//                 SELF  : one warp per (window, head): KV-cache append + attention over <= 448 positions
//                 CROSS : one 128-thread group per (window, head): 2 x 192 KB streamed, fp32 softmax in smem;
//                         the two groups of a CTA run out of phase so one's loads cover the other's softmax
// Activations written in one phase and read in another go through L2 (ld.global.cg / st.global) -- L1 is not
// coherent across SMs; weights use the read-only path.
#include "kernels.h"

namespace wjb {

constexpr int kMegaThreads = 256;

struct MegaSmem {
    union {
        float red[4][64][40];                 // GEMM partial tiles, up to 5 column tiles (40 KB)
        struct {
            float sc[2][1536];                // cross-attention scores per group
            float red2[2][8];
            float osum[2][4][64];
        } cross;
        float probs[8][448];                  // self-attention, per warp
    };
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ldcg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

// All CTAs of the grid are co-resident (grid <= #SMs, one CTA per SM by resource use).
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& target, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += nblocks;
        __threadfence();
        atomicAdd(bar, 1u);
        const long long t0 = clock64();
        while (ld_acquire_u32(bar) < target) {
            if (clock64() - t0 > 4000000000LL) {
                printf("wjb: grid barrier timed out (block %d, target %u, seen %u)\n", blockIdx.x, target, ld_acquire_u32(bar));
                __trap();
            }
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct GemmPhase {
    const __half* A;      // [B][lda]  (written earlier in this kernel: read through L2)
    int lda;
    const __half* W;      // [N][K]
    const __half* bias;   // [N] or null
    const __half* residual;  // [B][ld_out] or null (may alias out)
    __half* out;          // [B][ld_out]
    int N, K, ld_out, flags;
};

// NT tiles (8 columns each) at once.  A warp iteration covers U k-blocks per tile (U = 2 for NT <= 2, else 1); weights AND activations of iteration it+1 are issued before the
// MMAs of iteration it (register double buffering).
template <int NT>
__device__ __noinline__ void gemm_tiles(const GemmPhase& g, int B, const int (&tile)[5], MegaSmem& sm) {
    constexpr int U = (NT <= 2) ? 2 : 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gq = lane >> 2, t4 = lane & 3;
    const int nkb = g.K / 32;
    float acc[4][NT][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    const __half* wrow[NT];
    bool n_ok[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int n = tile[nt] * 8 + gq;
        n_ok[nt] = n < g.N;
        wrow[nt] = g.W + (size_t)(n_ok[nt] ? n : 0) * g.K + t4 * 8;
    }
    const __half* arow[8];
    bool a_ok[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = (i >> 1) * 16 + gq + 8 * (i & 1);
        a_ok[i] = r < B;
        arow[i] = g.A + (size_t)(a_ok[i] ? r : 0) * g.lda + t4 * 8;
    }
    // iteration it of warp w covers k-blocks (it * 8 + w) * U + u, u < U
    uint4 wv[NT][U], wn[NT][U], xv[U][8], xn[U][8];
    auto load_it = [&](uint4 (&wd)[NT][U], uint4 (&xd)[U][8], int it) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int kb = (it * 8 + warp) * U + u;
            const bool k_ok = kb < nkb;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                wd[nt][u] = (k_ok && n_ok[nt]) ? __ldg(reinterpret_cast<const uint4*>(wrow[nt] + (size_t)kb * 32)) : zero4;
#pragma unroll
            for (int i = 0; i < 8; ++i) xd[u][i] = (k_ok && a_ok[i]) ? ldcg16(arow[i] + (size_t)kb * 32) : zero4;
        }
    };
    const int n_it = (nkb + 8 * U - 1) / (8 * U);
    load_it(wv, xv, 0);
    for (int it = 0; it < n_it; ++it) {
        if (it + 1 < n_it) load_it(wn, xn, it + 1);
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                const uint4 x0 = xv[u][2 * mt], x1 = xv[u][2 * mt + 1];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    mma16816(acc[mt][nt], x0.x, x1.x, x0.y, x1.y, wv[nt][u].x, wv[nt][u].y);
                    mma16816(acc[mt][nt], x0.z, x1.z, x0.w, x1.w, wv[nt][u].z, wv[nt][u].w);
                }
            }
        }
        if (it + 1 < n_it) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) wv[nt][u] = wn[nt][u];
#pragma unroll
                for (int i = 0; i < 8; ++i) xv[u][i] = xn[u][i];
            }
        }
    }
    // fixed-order tree over the 8 warps (deterministic); the final tile lands in sm.red[0]
    auto store_acc = [&](int slot) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                *reinterpret_cast<float2*>(&sm.red[slot][mt * 16 + gq][nt * 8 + t4 * 2]) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
                *reinterpret_cast<float2*>(&sm.red[slot][mt * 16 + gq + 8][nt * 8 + t4 * 2]) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
            }
    };
    auto add_acc = [&](int slot) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float2 lo = *reinterpret_cast<const float2*>(&sm.red[slot][mt * 16 + gq][nt * 8 + t4 * 2]);
                const float2 hi = *reinterpret_cast<const float2*>(&sm.red[slot][mt * 16 + gq + 8][nt * 8 + t4 * 2]);
                acc[mt][nt][0] += lo.x;
                acc[mt][nt][1] += lo.y;
                acc[mt][nt][2] += hi.x;
                acc[mt][nt][3] += hi.y;
            }
    };
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
    if (warp == 0) {
        add_acc(0);
        store_acc(0);
    }
    __syncthreads();
    // epilogue by all 256 threads: thread -> (row m, column pair); loads first, then math, then stores
    constexpr int kPairs = NT * 4;              // column pairs per row
    constexpr int kOut = 64 * kPairs;           // outputs handled per CTA
    constexpr int kPer = (kOut + kMegaThreads - 1) / kMegaThreads;
    float v0[kPer], v1[kPer], r0[kPer], r1[kPer], b0[kPer], b1[kPer];
    size_t off[kPer];
    int ok[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
        const int idx = threadIdx.x + i * kMegaThreads;
        const int m = idx / kPairs, cp = idx % kPairs;
        const int nt = cp >> 2, c0 = tile[nt < NT ? nt : 0] * 8 + (cp & 3) * 2;
        ok[i] = 0;
        r0[i] = r1[i] = b0[i] = b1[i] = 0.f;
        off[i] = 0;
        if (idx < kOut && m < B && c0 < g.N) {
            ok[i] = (c0 + 1 < g.N) ? 2 : 1;
            off[i] = (size_t)m * g.ld_out + c0;
            const float2 t = *reinterpret_cast<const float2*>(&sm.red[0][m][nt * 8 + (cp & 3) * 2]);
            v0[i] = t.x;
            v1[i] = t.y;
            if (g.bias) {
                b0[i] = __half2float(__ldg(g.bias + c0));
                if (ok[i] == 2) b1[i] = __half2float(__ldg(g.bias + c0 + 1));
            }
            if (g.residual) {
                r0[i] = __half2float(__ldcg(g.residual + off[i]));
                if (ok[i] == 2) r1[i] = __half2float(__ldcg(g.residual + off[i] + 1));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
        if (!ok[i]) continue;
        float a0 = round_f16(v0[i] + b0[i]), a1 = round_f16(v1[i] + b1[i]);
        if (g.flags & GEMM_GELU) {
            a0 = round_f16(gelu_erf(a0));
            a1 = round_f16(gelu_erf(a1));
        }
        if (g.residual) {
            a0 += r0[i];
            a1 += r1[i];
        }
        if (ok[i] == 2 && ((off[i] & 1) == 0)) {
            *reinterpret_cast<__half2*>(g.out + off[i]) = __floats2half2_rn(a0, a1);
        } else {
            g.out[off[i]] = __float2half_rn(a0);
            if (ok[i] == 2) g.out[off[i] + 1] = __float2half_rn(a1);
        }
    }
    __syncthreads();  // sm.red is reused by the next chunk
}

__device__ void gemm_phase(const GemmPhase& g, int B, MegaSmem& sm) {
    const int n_tiles = (g.N + 7) / 8;
    const int grid = gridDim.x;
    // this CTA owns tiles blockIdx.x, blockIdx.x + grid, ...; up to 5 at once
    int t = blockIdx.x;
    while (t < n_tiles) {
        int tile[5] = {0, 0, 0, 0, 0};
        int cnt = 0;
        const int remaining = (n_tiles - t + grid - 1) / grid;
        const int take = remaining >= 8 ? 4 : (remaining > 5 ? (remaining + 1) / 2 : remaining);  // avoid a tiny last chunk
        for (; cnt < take && cnt < 5 && t < n_tiles; ++cnt, t += grid) tile[cnt] = t;
        switch (cnt) {
            case 5: gemm_tiles<5>(g, B, tile, sm); break;
            case 4: gemm_tiles<4>(g, B, tile, sm); break;
            case 3: gemm_tiles<3>(g, B, tile, sm); break;
            case 2: gemm_tiles<2>(g, B, tile, sm); break;
            default: gemm_tiles<1>(g, B, tile, sm); break;
        }
    }
}

// LayerNorm of row r by one warp: two-pass fp32 statistics, fp16 out (matches layernorm_kernel).
__device__ __forceinline__ void ln_row(const __half* x, const __half* gamma, const __half* beta, __half* out, int n) {
    const int lane = threadIdx.x & 31;
    const int nvec = n >> 3;
    float v[5][8];  // n <= 1280
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            const uint4 u = ldcg16(x + (size_t)idx * 8);
            const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h2[j]);
                v[i][2 * j] = f.x;
                v[i][2 * j + 1] = f.y;
                sum += f.x + f.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
        }
    }
    sum = warp_sum(sum);
    const float mean = sum / n;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i)
        if (lane + i * 32 < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = v[i][j] - mean;
                sq += d * d;
            }
        }
    sq = warp_sum(sq);
    const float rstd = rsqrtf(sq / n + 1e-5f);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            const uint4 gq = __ldg(reinterpret_cast<const uint4*>(gamma) + idx), bq = __ldg(reinterpret_cast<const uint4*>(beta) + idx);
            const __half2* gh = reinterpret_cast<const __half2*>(&gq);
            const __half2* bh = reinterpret_cast<const __half2*>(&bq);
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 gf = __half22float2(gh[j]), bf = __half22float2(bh[j]);
                oh[j] = __floats2half2_rn((v[i][2 * j] - mean) * rstd * gf.x + bf.x, (v[i][2 * j + 1] - mean) * rstd * gf.y + bf.y);
            }
            *reinterpret_cast<uint4*>(out + (size_t)idx * 8) = o;
        }
    }
}

__device__ __noinline__ void ln_phase(const __half* x, const __half* gamma, const __half* beta, __half* out, int B, int n) {
    const int warp = threadIdx.x >> 5;
    // row r -> CTA r % grid, warp (r / grid) % 8
    for (int r = blockIdx.x + gridDim.x * warp; r < B; r += gridDim.x * 8) ln_row(x + (size_t)r * n, gamma, beta, out + (size_t)r * n, n);
}

// Self-attention of one (b, h) by one warp (same arithmetic as attn_dec_self_kernel).
__device__ __noinline__ void self_attn_item(const __half* qkv_row, __half* kc, __half* vc, __half* out_row, int pos, float* probs) {
    const int lane = threadIdx.x & 31;
    // qkv_row points at this head's q; the caller has already appended this token's k, v to kc / vc
    float q[64];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(qkv_row);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 u = __ldcg(qp + i);
            const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h2[j]);
                q[i * 8 + 2 * j] = f.x;
                q[i * 8 + 2 * j + 1] = f.y;
            }
        }
    }
    float mx = -INFINITY;
    for (int p = lane; p <= pos; p += 32) {
        const uint4* kp = reinterpret_cast<const uint4*>(kc + (size_t)p * 64);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 u = __ldcg(kp + i);
            const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h2[j]);
                s = fmaf(q[i * 8 + 2 * j], f.x, s);
                s = fmaf(q[i * 8 + 2 * j + 1], f.y, s);
            }
        }
        s *= 0.125f;
        probs[p] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int p = lane; p <= pos; p += 32) {
        const float e = __expf(probs[p] - mx);
        probs[p] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int p = 0; p <= pos; ++p) {
        const float w = round_f16(probs[p] * inv);
        const float2 v = __half22float2(__ldcg(reinterpret_cast<const __half2*>(vc + (size_t)p * 64) + lane));
        o0 = fmaf(w, v.x, o0);
        o1 = fmaf(w, v.y, o1);
    }
    reinterpret_cast<__half2*>(out_row)[lane] = __floats2half2_rn(o0, o1);
    __syncwarp();
}

// Cross-attention of one (b, h) by a 128-thread group `grp` (named barrier 1 + grp).
__device__ __noinline__ void cross_attn_item(const __half* qh, const __half* Kh, const __half* Vh, __half* oh, int T, int grp, MegaSmem& sm) {
    const int tid = threadIdx.x & 127, lane = tid & 31, warp = tid >> 5;
    float* sc = sm.cross.sc[grp];
    float* red = sm.cross.red2[grp];
    auto gsync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory"); };
    const uint4* K = reinterpret_cast<const uint4*>(Kh);
    const uint4* V = reinterpret_cast<const uint4*>(Vh);
    const int chunk = tid & 7, slot = tid >> 3;
    float qf[8];
    {
        const uint4 u = __ldcg(reinterpret_cast<const uint4*>(qh) + chunk);
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h2[j]);
            qf[2 * j] = f.x;
            qf[2 * j + 1] = f.y;
        }
    }
    constexpr int UN = 16;  // 16-byte loads in flight per thread
    for (int t0 = 0; t0 < T; t0 += 16 * UN) {
        uint4 u[UN];
#pragma unroll
        for (int r = 0; r < UN; ++r) {
            const int t = t0 + r * 16 + slot;
            u[r] = (t < T) ? __ldg(K + (size_t)t * 8 + chunk) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int r = 0; r < UN; ++r) {
            const __half2* h2 = reinterpret_cast<const __half2*>(&u[r]);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h2[j]);
                s = fmaf(qf[2 * j], f.x, s);
                s = fmaf(qf[2 * j + 1], f.y, s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            const int t = t0 + r * 16 + slot;
            if (chunk == 0 && t < T) sc[t] = s * 0.125f;
        }
    }
    gsync();
    float mx = -INFINITY;
    for (int t = tid; t < T; t += 128) mx = fmaxf(mx, sc[t]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    gsync();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    float sum = 0.f;
    for (int t = tid; t < T; t += 128) {
        const float e = __expf(sc[t] - mx);
        sc[t] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[4 + warp] = sum;
    gsync();
    const float inv = 1.0f / (red[4] + red[5] + red[6] + red[7]);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int t0 = 0; t0 < T; t0 += 16 * UN) {
        uint4 u[UN];
#pragma unroll
        for (int r = 0; r < UN; ++r) {
            const int t = t0 + r * 16 + slot;
            u[r] = (t < T) ? __ldg(V + (size_t)t * 8 + chunk) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int r = 0; r < UN; ++r) {
            const int t = t0 + r * 16 + slot;
            const float w = (t < T) ? round_f16(sc[t] * inv) : 0.f;
            const __half2* h2 = reinterpret_cast<const __half2*>(&u[r]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h2[j]);
                acc[2 * j] = fmaf(w, f.x, acc[2 * j]);
                acc[2 * j + 1] = fmaf(w, f.y, acc[2 * j + 1]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
    }
    if (lane < 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sm.cross.osum[grp][warp][lane * 8 + j] = acc[j];
    }
    gsync();
    if (tid < 64) {
        const float v = sm.cross.osum[grp][0][tid] + sm.cross.osum[grp][1][tid] + sm.cross.osum[grp][2][tid] + sm.cross.osum[grp][3][tid];
        oh[tid] = __float2half_rn(v);
    }
    gsync();  // sc / osum reused by this group's next item
}

struct MegaArgs {
    const MegaLayer* layers;
    int n_layer;
    const __half *emb, *pos, *lnf_g, *lnf_b;
    int B, n, H, T, n_ctx, n_vocab, logits_stride;
    __half *x, *h, *qkv, *q, *a, *mlp, *logits;
    __half* self_kv;        // [layer][B][2H][n_ctx][64]
    const __half* cross_kv; // [layer][B][2H][T][64]
    const int* tokens;
    int tokens_stride;
    const DecodeCtl* ctl;
    const unsigned char* done;
    unsigned* bar;
    unsigned long long* prof;  // optional: globaltimer at every barrier exit of CTA 0 (debug)
};

__global__ void __launch_bounds__(kMegaThreads, 1) decode_mega_kernel(const MegaArgs a) {
    __shared__ MegaSmem sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grid = gridDim.x;
    const int B = a.B, n = a.n, H = a.H;
    const int step = a.ctl->step;
    unsigned target = 0;
    int prof_i = 0;
    auto stamp = [&]() {
        if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            a.prof[prof_i++] = t;
        }
    };
    stamp();

    // ---- embedding: row b -> CTA b % grid (warp 0..): x = half(emb[tok] + pos[step])
    for (int b = blockIdx.x + grid * warp; b < B; b += grid * 8) {
        const int tok = a.tokens[(size_t)b * a.tokens_stride + step];
        const __half2* e = reinterpret_cast<const __half2*>(a.emb + (size_t)tok * n);
        const __half2* p = reinterpret_cast<const __half2*>(a.pos + (size_t)step * n);
        __half2* o = reinterpret_cast<__half2*>(a.x + (size_t)b * n);
        for (int i = lane; i < n / 2; i += 32) {
            const float2 u = __half22float2(__ldg(e + i)), c = __half22float2(__ldg(p + i));
            o[i] = __floats2half2_rn(u.x + c.x, u.y + c.y);
        }
    }
    grid_barrier(a.bar, target, grid);
    stamp();

    const size_t self_row = (size_t)2 * H * a.n_ctx * 64, cross_row = (size_t)2 * H * a.T * 64;
    for (int l = 0; l < a.n_layer; ++l) {
        const MegaLayer& L = a.layers[l];
        __half* skv = a.self_kv + (size_t)l * B * self_row;
        const __half* ckv = a.cross_kv + (size_t)l * B * cross_row;
        // ---- self-attention block
        ln_phase(a.x, L.ln1_g, L.ln1_b, a.h, B, n);
        grid_barrier(a.bar, target, grid);
    stamp();
        {
            GemmPhase g{a.h, n, L.qkv_w, L.qkv_b, nullptr, a.qkv, 3 * n, n, 3 * n, 0};
            gemm_phase(g, B, sm);
        }
        grid_barrier(a.bar, target, grid);
    stamp();
        for (int item = blockIdx.x * 8 + warp; item < B * H; item += grid * 8) {
            const int b = item / H, hh = item % H;
            if (a.done[b]) continue;
            const __half* row = a.qkv + (size_t)b * 3 * n;
            __half* kc = skv + (size_t)b * self_row + (size_t)hh * a.n_ctx * 64;
            __half* vc = skv + (size_t)b * self_row + (size_t)(H + hh) * a.n_ctx * 64;
            reinterpret_cast<__half2*>(kc + (size_t)step * 64)[lane] = __ldcg(reinterpret_cast<const __half2*>(row + n + hh * 64) + lane);
            reinterpret_cast<__half2*>(vc + (size_t)step * 64)[lane] = __ldcg(reinterpret_cast<const __half2*>(row + 2 * n + hh * 64) + lane);
            __syncwarp();
            self_attn_item(row + hh * 64, kc, vc, a.a + (size_t)b * n + hh * 64, step, sm.probs[warp]);
        }
        grid_barrier(a.bar, target, grid);
    stamp();
        {
            GemmPhase g{a.a, n, L.out_w, L.out_b, a.x, a.x, n, n, n, 0};
            gemm_phase(g, B, sm);
        }
        grid_barrier(a.bar, target, grid);
    stamp();
        // ---- cross-attention block
        ln_phase(a.x, L.ln2_g, L.ln2_b, a.h, B, n);
        grid_barrier(a.bar, target, grid);
    stamp();
        {
            GemmPhase g{a.h, n, L.cq_w, L.cq_b, nullptr, a.q, n, n, n, 0};
            gemm_phase(g, B, sm);
        }
        grid_barrier(a.bar, target, grid);
    stamp();
        {
            const int grp = threadIdx.x >> 7;
            for (int item = blockIdx.x * 2 + grp; item < B * H; item += grid * 2) {
                const int b = item / H, hh = item % H;
                if (a.done[b]) continue;
                cross_attn_item(a.q + (size_t)b * n + hh * 64, ckv + (size_t)b * cross_row + (size_t)hh * a.T * 64,
                                ckv + (size_t)b * cross_row + (size_t)(H + hh) * a.T * 64, a.a + (size_t)b * n + hh * 64, a.T, grp, sm);
            }
        }
        grid_barrier(a.bar, target, grid);
    stamp();
        {
            GemmPhase g{a.a, n, L.cout_w, L.cout_b, a.x, a.x, n, n, n, 0};
            gemm_phase(g, B, sm);
        }
        grid_barrier(a.bar, target, grid);
    stamp();
        // ---- MLP block
        ln_phase(a.x, L.ln3_g, L.ln3_b, a.h, B, n);
        grid_barrier(a.bar, target, grid);
    stamp();
        {
            GemmPhase g{a.h, n, L.fc1_w, L.fc1_b, nullptr, a.mlp, 4 * n, n, 4 * n, GEMM_GELU};
            gemm_phase(g, B, sm);
        }
        grid_barrier(a.bar, target, grid);
    stamp();
        {
            GemmPhase g{a.mlp, 4 * n, L.fc2_w, L.fc2_b, a.x, a.x, n, 4 * n, n, 0};
            gemm_phase(g, B, sm);
        }
        grid_barrier(a.bar, target, grid);
    stamp();
    }
    ln_phase(a.x, a.lnf_g, a.lnf_b, a.h, B, n);
    grid_barrier(a.bar, target, grid);
    stamp();
    {
        GemmPhase g{a.h, n, a.emb, nullptr, nullptr, a.logits, a.n_vocab, n, a.logits_stride, 0};
        gemm_phase(g, B, sm);
    }
    stamp();
}

int launch_decode_mega(const MegaLaunch& m, cudaStream_t s) {
    if (m.B < 1 || m.B > 64) return set_error("decode_mega: batch %d out of range", m.B);
    if (m.n % 32 || m.n > 1280 || m.T > 1536 || m.n_ctx > 448) return set_error("decode_mega: unsupported dims");
    MegaArgs a;
    a.layers = m.layers;
    a.n_layer = m.n_layer;
    a.emb = m.emb;
    a.pos = m.pos;
    a.lnf_g = m.lnf_g;
    a.lnf_b = m.lnf_b;
    a.B = m.B;
    a.n = m.n;
    a.H = m.H;
    a.T = m.T;
    a.n_ctx = m.n_ctx;
    a.n_vocab = m.n_vocab;
    a.logits_stride = m.logits_stride;
    a.x = m.x;
    a.h = m.h;
    a.qkv = m.qkv;
    a.q = m.q;
    a.a = m.a;
    a.mlp = m.mlp;
    a.logits = m.logits;
    a.self_kv = m.self_kv;
    a.cross_kv = m.cross_kv;
    a.tokens = m.tokens;
    a.tokens_stride = m.tokens_stride;
    a.ctl = m.ctl;
    a.done = m.done;
    a.bar = m.bar;
    a.prof = m.prof;
    cudaError_t e = cudaMemsetAsync(m.bar, 0, sizeof(unsigned), s);
    if (e != cudaSuccess) return set_error("decode_mega memset: %s", cudaGetErrorString(e));
    decode_mega_kernel<<<sm_count(), kMegaThreads, 0, s>>>(a);
    WJB_CHECK_LAUNCH("decode_mega");
    return 0;
}

}  // namespace wjb
