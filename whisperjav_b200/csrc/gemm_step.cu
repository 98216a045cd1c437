// Decoder-step GEMM: out[r][n] = epilogue(sum_k A[r][k] W[n][k]) for <= 128 activation rows (one batch of decoder
// rows), fp16 in, fp32 accumulate -- the Linear layers of openai-whisper model.py::ResidualAttentionBlock at one
// autoregressive position.
//
// Such a GEMM is pure weight streaming, but one launch of the general tcgen05 kernel is bound by things that have
// nothing to do with HBM (measured with the timeline hook, scripts/gemm_trace.py, and scripts/mma_rate.cu):
//   * tcgen05.mma (M = 128, K = 16) costs 61 / 74 / 136 clk at N = 64 / 128 / 256 whether A comes from shared memory or
//     from TMEM, but a tcgen05.commit between MMAs drains the pipe (~300 clk): with one commit per k block (needed to
//     recycle a pipeline stage) a narrow tile runs at 300 ns per 64 columns of K, and K = 1280 costs 6 us per CTA,
//   * an output-tile-only decomposition gives 20-80 CTAs that each re-read the whole activation block.
// So here
//   * a thread-block cluster of S CTAs shares one output tile (128 x BN) and slices K; the activation slice of a CTA
//     arrives with one or two TMA loads (3-D box: 64 columns x rows x k blocks) and the weight slice is resident as a
//     whole whenever it fits, so the k loop issues its MMAs back to back with a single commit at the end,
//   * the fp32 partial tiles meet through distributed shared memory: CTA r owns BN/S columns, every CTA writes its
//     partials of those columns into r's receive slab (st.shared::cluster, a warp writes 512 contiguous bytes), one
//     cluster barrier, and r adds the S slabs in rank order (deterministic) and runs the epilogue (bias, GELU,
//     residual) on its columns,
//   * optionally the LayerNorm in front of the Linear (model.py::ResidualAttentionBlock: attn_ln / cross_attn_ln / mlp_ln)
//     runs inside: every CTA holds a K slice of all rows, so the cluster adds per-row partial sums through distributed
//     shared memory (two passes: mean, then sum of squared deviations, fp32, rank order), normalises its tile in place
//     and hands it to the MMA thread -- one launch and one global round trip less per LayerNorm,
//   * ... or without any exchange (decode-step option 1 of wjb_debug_set_decode_flags; correct, tested, measured no faster than the
//     separate LayerNorm launch and therefore off by default -- DESIGN.md section 4): the GEMM that PRODUCES the residual stream
//     (out / cross-out / fc2) adds
//     per-row sum and sum of squares of the fp16 values it stores to a fixed-point accumulator (64-bit integer atomics: the
//     order of the adds cannot change the result), and the GEMM behind the LayerNorm reads two integers per row, normalises
//     its tile in shared memory with the reference's rounding points (fp32 statistics of the fp16 row, fp16-rounded output)
//     and starts its MMAs: 97 LayerNorm launches per decoder step become 1,
//   * weights are constants, so their TMA loads are issued BEFORE griddepcontrol.wait: under programmatic dependent
//     launch the CTAs of this GEMM are resident while the previous kernels of the step still run and the weight
//     slices are already in shared memory when the activations become valid.
#include "kernels.h"

namespace wjb {

constexpr int kStepThreads = 192;  // warp 0 TMA, warp 1 MMA + TMEM owner, warps 2..5 epilogue (TMEM lane quadrant = warp % 4)
constexpr int kStepBlockK = 64;    // 128 B of fp16 per row: one SWIZZLE_128B atom
constexpr int kStepMaxKb = 20;     // k blocks per CTA
constexpr int kStepMaxStages = 20;
constexpr int kStepSmemBudget = 216 * 1024;

struct StepDev {
    const __half* bias;
    const __half* residual;
    __half* out;
    long long out_row_stride;
    int rows, N, K, flags;
    int a_tile;       // bytes of one activation k block in shared memory (TMA box rows x 128)
    int a_chunks;     // TMA loads the activation slice arrives in (1 or 2)
    int kpc;          // k blocks per such load
    int w_stages;
    int recv_rows;    // rows per receive slab (rows rounded up to 32)
    int w_early;      // 1 = weights may be fetched before the previous kernel has finished
    const __half* ln_g;  // LayerNorm over the K columns of every activation row, applied to the tile in shared memory before the
    const __half* ln_b;  // MMAs (null = activations are used as they are)
    const long long* ln_stats;  // [rows][2] fixed-point (2^20) sum / sum of squares of every activation row: no cluster exchange needed
    long long* out_stats;       // [rows][2] accumulators this GEMM adds the statistics of its OUTPUT rows to (null = none)
    const void* pf_ptr;         // constant region (the next Linear's weights) this grid pulls into L2 while it runs (null = none)
    unsigned long long pf_bytes;
    unsigned long long* trace;
};

#define WJB_STEP_TRACE(k)                                               \
    do {                                                                \
        if (trace_row) {                                                \
            unsigned long long gt_;                                     \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));     \
            trace_row[2 * (k)] = clock64();                             \
            trace_row[2 * (k) + 1] = gt_;                               \
        }                                                               \
    } while (0)

WJB_DEVINL uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
WJB_DEVINL uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
WJB_DEVINL void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
WJB_DEVINL void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
WJB_DEVINL uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
WJB_DEVINL void st_cluster_f4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

WJB_DEVINL void st_cluster_f1(uint32_t addr, float a) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory"); }

template <int BN>
__global__ void __launch_bounds__(kStepThreads, 1)
gemm_step_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const StepDev p) {
    constexpr int kTmemCols = BN < 32 ? 32 : BN;
    constexpr int kBBytes = BN * 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w_tiles = smem;                          // [w_stages][BN][128 B]              SWIZZLE_128B
    uint8_t* a_tiles = smem + p.w_stages * kBBytes;   // [a_chunks * kpc][box rows][128 B]  SWIZZLE_128B, + 8 KB the MMA may over-read
    float* recv = reinterpret_cast<float*>(a_tiles + p.a_chunks * p.kpc * p.a_tile + 8192);  // [S][BN / S / 4][recv_rows] float4
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(recv) + (size_t)p.recv_rows * BN * 4);
    uint64_t* w_full = bars;
    uint64_t* w_empty = w_full + kStepMaxStages;
    uint64_t* a_full = w_empty + kStepMaxStages;  // one per activation load
    uint64_t* tfull_bar = a_full + 2;
    uint64_t* a_ready = tfull_bar + 1;  // LayerNorm variant: the normalised tile is in place
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);
    float* lnbuf = reinterpret_cast<float*>(bars + 2 * kStepMaxStages + 8);  // [2 passes][S][128 rows] partial sums
    const bool ln = p.ln_g != nullptr;
    const bool ln_x = ln && p.ln_stats == nullptr;  // LayerNorm with the statistics exchanged inside the cluster

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t S = cluster_nctarank(), rank = cluster_ctarank();
    const int nt = blockIdx.x / S;  // output tile (columns nt * BN ...)
    const int num_kb_total = p.K / kStepBlockK;
    const int kb0 = num_kb_total * (int)rank / (int)S, kb1 = num_kb_total * ((int)rank + 1) / (int)S;
    const int nkb = kb1 - kb0;

    __shared__ unsigned long long* trace_sh;
    unsigned long long* trace_row = nullptr;
    if (threadIdx.x == 0) {
        if (p.trace && blockIdx.x == 0) {
            const unsigned long long launch_idx = atomicAdd(p.trace, 1ull);
            trace_row = p.trace + 32 + (launch_idx % 512) * 32;  // ring of the last 512 launches
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            trace_row[30] = smid;
            trace_row[29] = launch_idx;
            trace_row[28] = ((unsigned long long)p.N << 32) | (unsigned)p.K;
            WJB_STEP_TRACE(0);
        }
        trace_sh = trace_row;
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < p.w_stages; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&w_empty[s], 1);
        }
        mbar_init(&a_full[0], 1);
        mbar_init(&a_full[1], 1);
        mbar_init(tfull_bar, 1);
        mbar_init(a_ready, 4);
        fence_barrier_init();
        // weight slices first: they neither depend on the previous kernel nor on TMEM
        const int pre = p.w_early ? (nkb < p.w_stages ? nkb : p.w_stages) : 0;
        for (int i = 0; i < pre; ++i) {
            mbar_arrive_expect_tx(&w_full[i], kBBytes);
            tma_load_2d(w_tiles + i * kBBytes, &tmB, &w_full[i], (kb0 + i) * kStepBlockK, nt * BN);
        }
        l2_prefetch_share(p.pf_ptr, p.pf_bytes, blockIdx.x, gridDim.x);
        WJB_STEP_TRACE(1);
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_arrive();  // "this CTA is running": matched by the wait in front of the first remote store
    if (ln_x) cluster_wait();  // ... which in the exchanging LayerNorm variant is the exchange of partial sums, so wait right away
    const uint32_t tmem_base = *tmem_slot;
    if (lane == 0) trace_row = trace_sh;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const int pre = p.w_early ? (nkb < p.w_stages ? nkb : p.w_stages) : 0;
            asm volatile("griddepcontrol.wait;" ::: "memory");
            WJB_STEP_TRACE(2);
            for (int c = 0; c < p.a_chunks; ++c) {  // the whole activation slice: one 3-D box per load
                mbar_arrive_expect_tx(&a_full[c], (uint32_t)(p.kpc * p.a_tile));
                tma_load_3d(a_tiles + c * p.kpc * p.a_tile, &tmA, &a_full[c], 0, 0, kb0 + c * p.kpc);
            }
            WJB_STEP_TRACE(3);
        }
        __syncwarp();
        if (ln_x) {  // every thread of the cluster takes part in the two partial-sum exchanges
            cluster_arrive();
            cluster_wait();
            cluster_arrive();
            cluster_wait();
        }
        if (lane == 0) {
            const int pre = p.w_early ? (nkb < p.w_stages ? nkb : p.w_stages) : 0;
            int ws = pre % p.w_stages;
            uint32_t wph = (pre / p.w_stages) & 1;
            for (int i = pre; i < nkb; ++i) {
                mbar_wait(&w_empty[ws], wph ^ 1);
                mbar_arrive_expect_tx(&w_full[ws], kBBytes);
                tma_load_2d(w_tiles + ws * kBBytes, &tmB, &w_full[ws], (kb0 + i) * kStepBlockK, nt * BN);
                if (++ws == p.w_stages) {
                    ws = 0;
                    wph ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (ln_x) {
            cluster_arrive();
            cluster_wait();
            cluster_arrive();
            cluster_wait();
        }
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(128, BN, 0, 0);
            if (ln) {  // activation barriers first, then the normalised tile
                mbar_wait(&a_full[0], 0);
                if (p.a_chunks > 1) mbar_wait(&a_full[1], 0);
                mbar_wait(a_ready, 0);
            }
            if (p.w_stages >= nkb) {
                // the whole weight slice is resident: collect every barrier first, then issue the MMAs back to back
                // (a try_wait between groups of MMAs lets the tensor pipe run dry, scripts/mma_rate.cu)
                for (int i = 0; i < nkb; ++i) mbar_wait(&w_full[i], 0);
                mbar_wait(&a_full[0], 0);
                tc_fence_after();
                WJB_STEP_TRACE(4);
                for (int i = 0; i < nkb; ++i) {
                    if (i == p.kpc) {
                        mbar_wait(&a_full[1], 0);
                        tc_fence_after();
                    }
                    if (i >= 1 && i <= 5) WJB_STEP_TRACE(8 + i);
                    const uint64_t da = make_smem_desc(smem_u32(a_tiles + i * p.a_tile), 16, 1024, kLayoutSW128);
                    const uint64_t db = make_smem_desc(smem_u32(w_tiles + i * kBBytes), 16, 1024, kLayoutSW128);
#pragma unroll
                    for (int k = 0; k < kStepBlockK / 16; ++k) umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (i | k) != 0);
                }
            } else {
                int ws = 0;
                uint32_t wph = 0;
                for (int i = 0; i < nkb; ++i) {
                    mbar_wait(&w_full[ws], wph);
                    if (i == 0) mbar_wait(&a_full[0], 0);
                    if (i == p.kpc) mbar_wait(&a_full[1], 0);
                    tc_fence_after();
                    if (i == 0) WJB_STEP_TRACE(4);
                    if (i >= 1 && i <= 5) WJB_STEP_TRACE(8 + i);
                    // rows <= 64: the tile is 8 KB and the MMA's upper 64 rows read the next tile; their accumulator lanes are never used
                    const uint64_t da = make_smem_desc(smem_u32(a_tiles + i * p.a_tile), 16, 1024, kLayoutSW128);
                    const uint64_t db = make_smem_desc(smem_u32(w_tiles + ws * kBBytes), 16, 1024, kLayoutSW128);
#pragma unroll
                    for (int k = 0; k < kStepBlockK / 16; ++k) umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (i | k) != 0);
                    // a commit drains the tensor pipe before the next MMA starts (~300 clk, scripts/mma_rate.cu): only pay for it
                    // when the weight stage is going to be refilled
                    if (i + p.w_stages < nkb) umma_commit(&w_empty[ws]);
                    if (++ws == p.w_stages) {
                        ws = 0;
                        wph ^= 1;
                    }
                }
            }
            umma_commit(tfull_bar);
            WJB_STEP_TRACE(5);
        }
    }
    __syncwarp();

    const int w = BN / (int)S;  // columns each CTA of the cluster finishes
    if (warp >= 2) {
        const int q = warp & 3;  // TMEM lane quadrant
        const int row = q * 32 + lane;
        if (ln && !ln_x) {  // gamma / beta of this CTA's K slice are constants: stage them in shared memory ahead of the dependency wait
            const int t = threadIdx.x - 64;
            uint4* gsm = reinterpret_cast<uint4*>(lnbuf + 512);  // [nkb * 8] gamma, then [nkb * 8] beta (16-byte pieces of 8 halfs)
            const uint4* gg = reinterpret_cast<const uint4*>(p.ln_g + (long long)kb0 * kStepBlockK);
            const uint4* gb = reinterpret_cast<const uint4*>(p.ln_b + (long long)kb0 * kStepBlockK);
            for (int idx = t; idx < nkb * 8; idx += 128) {
                gsm[idx] = __ldg(gg + idx);
                gsm[nkb * 8 + idx] = __ldg(gb + idx);
            }
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");  // the epilogue reads the residual and overwrites `out`
        if (ln && !ln_x) {
            // ===================== LayerNorm from the producer's row statistics, tile normalised in place =====================
            const int t = threadIdx.x - 64;
            float2* rowstat = reinterpret_cast<float2*>(lnbuf);  // [rows] (mean, rstd)
            const uint4* gsm = reinterpret_cast<const uint4*>(lnbuf + 512);
            if (t < p.rows) {
                const long long sfx = p.ln_stats[2 * t], qfx = p.ln_stats[2 * t + 1];
                const double inv = 1.0 / ((double)p.K * 1048576.0);
                const double mean = (double)sfx * inv;
                double var = (double)qfx * inv - mean * mean;
                var = var > 0.0 ? var : 0.0;
                rowstat[t] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait_warp(&a_full[0], 0);
            if (p.a_chunks > 1) mbar_wait_warp(&a_full[1], 0);
            // thread -> 16-byte chunk c of every k block, rows (t >> 3) + 16 j: gamma / beta are loaded once per k block
            const int c = t & 7, r0 = t >> 3;
            for (int i = 0; i < nkb; ++i) {
                const uint4 g = gsm[i * 8 + c];
                const uint4 bb = gsm[nkb * 8 + i * 8 + c];
                const __half2* gh = reinterpret_cast<const __half2*>(&g);
                const __half2* bh = reinterpret_cast<const __half2*>(&bb);
                float2 gf[4], bf[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    gf[e] = __half22float2(gh[e]);
                    bf[e] = __half22float2(bh[e]);
                }
                for (int lrow = r0; lrow < p.rows; lrow += 16) {
                    const float2 st = rowstat[lrow];
                    uint4* px = reinterpret_cast<uint4*>(a_tiles + i * p.a_tile + lrow * 128 + ((c ^ (lrow & 7)) << 4));
                    uint4 u = *px;
                    __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = __half22float2(h2[e]);
                        h2[e] = __floats2half2_rn((f.x - st.x) * st.y * gf[e].x + bf[e].x, (f.y - st.x) * st.y * gf[e].y + bf[e].y);
                    }
                    *px = u;
                }
            }
            fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's reads of shared memory
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
        } else if (ln) {
            // ===================== LayerNorm of the activation tile, in place =====================
            // rows <= 64: two threads per row (4 of the 8 16-byte chunks of a k block each), else one thread per row
            const int t = threadIdx.x - 64;
            const int tpr = p.rows <= 64 ? 2 : 1;
            const int lrow = t / tpr, part = t % tpr;
            const bool active = lrow < p.rows;
            const int c0 = part * (8 / tpr), c1 = c0 + 8 / tpr;
            const uint32_t lnbuf_local = smem_u32(lnbuf);
            mbar_wait_warp(&a_full[0], 0);
            if (p.a_chunks > 1) mbar_wait_warp(&a_full[1], 0);
            const uint8_t* rbase = a_tiles + lrow * 128;
            float mean = 0.f, rstd = 0.f;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                float acc = 0.f;
                if (active) {
                    for (int i = 0; i < nkb; ++i) {
                        for (int j = c0; j < c1; ++j) {
                            const uint4 u = *reinterpret_cast<const uint4*>(rbase + i * p.a_tile + ((j ^ (lrow & 7)) << 4));
                            const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __half22float2(h2[e]);
                                if (pass == 0) {
                                    acc += f.x + f.y;
                                } else {
                                    const float d0 = f.x - mean, d1 = f.y - mean;
                                    acc += d0 * d0 + d1 * d1;
                                }
                            }
                        }
                    }
                }
                if (tpr == 2) acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                if (active && part == 0) {
                    const uint32_t off = ((uint32_t)(pass * 8 + (int)rank) * 128 + (uint32_t)lrow) * 4;
                    for (uint32_t dst = 0; dst < S; ++dst) st_cluster_f1(map_to_rank(lnbuf_local + off, dst), acc);
                }
                cluster_arrive();
                cluster_wait();
                float tot = 0.f;
                if (active)
                    for (uint32_t sr = 0; sr < S; ++sr) tot += lnbuf[(pass * 8 + sr) * 128 + lrow];
                if (pass == 0)
                    mean = tot / (float)p.K;
                else
                    rstd = rsqrtf(tot / (float)p.K + 1e-5f);
            }
            if (active) {
                for (int i = 0; i < nkb; ++i) {
                    for (int j = c0; j < c1; ++j) {
                        uint4* px = reinterpret_cast<uint4*>(const_cast<uint8_t*>(rbase) + i * p.a_tile + ((j ^ (lrow & 7)) << 4));
                        uint4 u = *px;
                        const long long kcol = (long long)(kb0 + i) * kStepBlockK + j * 8;
                        const uint4 g = __ldg(reinterpret_cast<const uint4*>(p.ln_g + kcol));
                        const uint4 bb = __ldg(reinterpret_cast<const uint4*>(p.ln_b + kcol));
                        __half2* h2 = reinterpret_cast<__half2*>(&u);
                        const __half2* gh = reinterpret_cast<const __half2*>(&g);
                        const __half2* bh = reinterpret_cast<const __half2*>(&bb);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = __half22float2(h2[e]), gf = __half22float2(gh[e]), bf = __half22float2(bh[e]);
                            h2[e] = __floats2half2_rn((f.x - mean) * rstd * gf.x + bf.x, (f.y - mean) * rstd * gf.y + bf.y);
                        }
                        *px = u;
                    }
                }
            }
            fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's reads of shared memory
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
        }
        // ===================== partial tiles -> owners' receive slabs =====================
        mbar_wait_warp(tfull_bar, 0);
        tc_fence_after();
        if (warp == 2) WJB_STEP_TRACE(6);
        if (!ln_x) cluster_wait();  // every CTA of the cluster is running: its shared memory may be written
        if (q * 32 < p.rows) {
            const uint32_t recv_local = smem_u32(recv);
            const uint32_t qpo = (uint32_t)w / 4;  // float4 groups per owner
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + c, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const uint32_t g = (uint32_t)(c + j) / 4;  // float4 group of the tile
                    const uint32_t owner = g / qpo, qd = g % qpo;
                    // slab [source rank][group][row]: the 32 lanes of a store write 512 contiguous bytes
                    const uint32_t off = ((rank * qpo + qd) * (uint32_t)p.recv_rows + (uint32_t)row) * 16;
                    st_cluster_f4(map_to_rank(recv_local + off, owner), __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                  __uint_as_float(r[j + 3]));
                }
            }
        }
        tc_fence_before();
    } else if (!ln_x) {
        cluster_wait();
    }
    cluster_arrive();
    cluster_wait();  // all partials of my columns have landed
    if (warp == 2) WJB_STEP_TRACE(7);

    // ===================== add the S slabs in rank order, epilogue, store =====================
    if (warp >= 2) {
        const int t = threadIdx.x - 64;  // 0..127
        const int qpo = w / 4;
        const int total = p.rows * qpo;
        const int col_base = nt * BN + (int)rank * w;
        const float4* slab = reinterpret_cast<const float4*>(recv);
        unsigned long long* rowacc = reinterpret_cast<unsigned long long*>(lnbuf);  // fixed-point partial sums of this CTA's columns
        // rows dividing 128 (the decode batch of 64): a thread meets the same row in every iteration, so it keeps its partial sums in
        // registers and the 128 / rows threads of a row meet through one shared-memory slot each; other row counts: shared atomics
        const bool fixed_row = p.out_stats && (128 % p.rows) == 0;
        long long my_sm = 0, my_sq = 0;
        if (p.out_stats && !fixed_row) {
            if (t < p.rows) rowacc[2 * t] = rowacc[2 * t + 1] = 0ull;
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        for (int e = t; e < total; e += 128) {
            const int row = e % p.rows, qd = e / p.rows;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (uint32_t s = 0; s < S; ++s) {
                const float4 v = slab[((size_t)s * qpo + qd) * p.recv_rows + row];
                acc.x += v.x;
                acc.y += v.y;
                acc.z += v.z;
                acc.w += v.w;
            }
            const int col = col_base + qd * 4;
            float v[4] = {acc.x, acc.y, acc.z, acc.w};
            if (p.bias) {
                const uint2 bb = __ldg(reinterpret_cast<const uint2*>(p.bias + col));
                const __half2* h2 = reinterpret_cast<const __half2*>(&bb);
                const float2 b0 = __half22float2(h2[0]), b1 = __half22float2(h2[1]);
                v[0] += b0.x;
                v[1] += b0.y;
                v[2] += b1.x;
                v[3] += b1.y;
            }
            if (p.flags & GEMM_GELU) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = gelu_erf_fast(round_f16(v[j]));
            }
            const long long off = (long long)row * p.out_row_stride + col;
            if (p.residual) {
                const uint2 rv = *reinterpret_cast<const uint2*>(p.residual + off);
                const __half2* rh = reinterpret_cast<const __half2*>(&rv);
                const float2 r0 = __half22float2(rh[0]), r1 = __half22float2(rh[1]);
                v[0] = round_f16(v[0]) + r0.x;  // the Linear output is an fp16 tensor before the residual add
                v[1] = round_f16(v[1]) + r0.y;
                v[2] = round_f16(v[2]) + r1.x;
                v[3] = round_f16(v[3]) + r1.y;
            }
            uint2 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
            oh[0] = __floats2half2_rn(v[0], v[1]);
            oh[1] = __floats2half2_rn(v[2], v[3]);
            *reinterpret_cast<uint2*>(p.out + off) = o;
            if (p.out_stats) {  // statistics of the values as stored (the next LayerNorm reads the fp16 tensor)
                // element by element in fixed point: x * 2^20 and x^2 * 2^20 are exact in fp32 for fp16 x, so the sums are exact up to
                // the 2^-21 rounding of each square -- the variance is then E[x^2] - mean^2 of exact numbers (fp64 in the consumer),
                // as accurate as a two-pass fp32 LayerNorm even when the row's mean dwarfs its spread
                const float2 f0 = __half22float2(oh[0]), f1 = __half22float2(oh[1]);
                const long long sm = __float2ll_rn(f0.x * 1048576.0f) + __float2ll_rn(f0.y * 1048576.0f) + __float2ll_rn(f1.x * 1048576.0f) +
                                     __float2ll_rn(f1.y * 1048576.0f);
                const long long sq = __float2ll_rn(f0.x * f0.x * 1048576.0f) + __float2ll_rn(f0.y * f0.y * 1048576.0f) +
                                     __float2ll_rn(f1.x * f1.x * 1048576.0f) + __float2ll_rn(f1.y * f1.y * 1048576.0f);
                if (fixed_row) {
                    my_sm += sm;
                    my_sq += sq;
                } else {
                    atomicAdd(&rowacc[2 * row], (unsigned long long)sm);
                    atomicAdd(&rowacc[2 * row + 1], (unsigned long long)sq);
                }
            }
        }
        if (p.out_stats) {
            if (fixed_row) {  // slot [t]: this thread's partial of row t % rows
                rowacc[2 * t] = (unsigned long long)my_sm;
                rowacc[2 * t + 1] = (unsigned long long)my_sq;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (t < p.rows) {
                unsigned long long a = rowacc[2 * t], b = rowacc[2 * t + 1];
                if (fixed_row)
                    for (int u = t + p.rows; u < 128; u += p.rows) {
                        a += rowacc[2 * u];
                        b += rowacc[2 * u + 1];
                    }
                atomicAdd(reinterpret_cast<unsigned long long*>(p.out_stats) + 2 * t, a);
                atomicAdd(reinterpret_cast<unsigned long long*>(p.out_stats) + 2 * t + 1, b);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) WJB_STEP_TRACE(8);
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

static int encode_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes, uint32_t box_inner,
                     uint32_t box_outer) {
    auto encode = get_tensor_map_encoder();
    if (!encode) return set_error("cuTensorMapEncodeTiled unavailable");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("gemm_step: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

static unsigned long long* g_step_trace = nullptr;
void gemm_step_set_trace(void* buf) { g_step_trace = reinterpret_cast<unsigned long long*>(buf); }

static int encode_a_3d(CUtensorMap* map, const void* base, int K, int rows, long long row_stride_halfs, int box_rows, int kpc) {
    auto encode = get_tensor_map_encoder();
    if (!encode) return set_error("cuTensorMapEncodeTiled unavailable");
    // [k block][row][64 halfs]: the k-block dimension is outermost in the box but has the smallest stride in memory
    cuuint64_t dims[3] = {(cuuint64_t)kStepBlockK, (cuuint64_t)rows, (cuuint64_t)(K / kStepBlockK)};
    cuuint64_t strides[2] = {(cuuint64_t)row_stride_halfs * 2, (cuuint64_t)kStepBlockK * 2};
    cuuint32_t box[3] = {(cuuint32_t)kStepBlockK, (cuuint32_t)box_rows, (cuuint32_t)kpc};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("gemm_step: cuTensorMapEncodeTiled (activations) failed (%d)", (int)r);
    return 0;
}

// shared-memory plan of one configuration; false when it does not fit
static bool plan_step(int rows, int K, int bn, int S, StepDev* d, size_t* smem_bytes) {
    const int num_kb = K / kStepBlockK;
    const int max_slice = (num_kb + S - 1) / S;
    if (max_slice > kStepMaxKb || num_kb < S) return false;
    const int box_rows = rows <= 64 ? 64 : 128;
    d->a_tile = box_rows * 128;
    d->a_chunks = max_slice >= 6 ? 2 : 1;
    d->kpc = (max_slice + d->a_chunks - 1) / d->a_chunks;
    d->recv_rows = (rows + 31) / 32 * 32;
    const int recv_bytes = d->recv_rows * bn * 4;
    const int a_bytes = d->a_chunks * d->kpc * d->a_tile + 8192;
    int ws = (kStepSmemBudget - 8192 /*LayerNorm partial sums*/ - recv_bytes - a_bytes) / (bn * 128);
    if (ws > max_slice) ws = max_slice;
    if (ws > kStepMaxStages) ws = kStepMaxStages;
    if (ws < 1 || (ws < 2 && max_slice >= 2)) return false;
    d->w_stages = ws;
    *smem_bytes = 1024 + (size_t)ws * bn * 128 + a_bytes + recv_bytes + (2 * kStepMaxStages + 8) * 8 + 2 * 8 * 128 * 4 + 64;
    return true;
}

template <int BN>
static int launch_step_bn(const StepGemmArgs& a, int S, cudaStream_t stream) {
    StepDev d;
    size_t smem = 0;
    if (!plan_step(a.rows, a.K, BN, S, &d, &smem)) return set_error("gemm_step: rows=%d K=%d BN=%d S=%d does not fit shared memory", a.rows, a.K, BN, S);
    d.bias = a.bias;
    d.residual = a.residual;
    d.out = a.out;
    d.out_row_stride = a.out_row_stride;
    d.rows = a.rows;
    d.N = a.N;
    d.K = a.K;
    d.flags = a.flags;
    d.w_early = a.w_constant ? 1 : 0;
    d.ln_g = a.ln_gamma;
    d.ln_b = a.ln_beta;
    d.ln_stats = a.ln_stats;
    d.out_stats = a.out_stats;
    d.pf_ptr = a.prefetch;
    d.pf_bytes = a.prefetch_bytes;
    d.trace = g_step_trace;
    CUtensorMap tmA, tmB;
    if (int e = encode_a_3d(&tmA, a.A, a.K, a.rows, a.a_row_stride, d.a_tile / 128, d.kpc)) return e;
    if (int e = encode_2d(&tmB, a.W, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.ldw * 2, kStepBlockK, (uint32_t)BN)) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((a.N / BN) * S);
    cfg.blockDim = dim3(kStepThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    int n = 0;
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = S;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
    if (pdl_enabled()) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_step_kernel<BN>, tmA, tmB, d);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("gemm_step launch (BN=%d S=%d smem=%zu): %s", BN, S, smem, cudaGetErrorString(e));
    return 0;
}

int gemm_step_init() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(gemm_step_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStepSmemBudget + 2048)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(gemm_step_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStepSmemBudget + 2048)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(gemm_step_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStepSmemBudget + 2048)) != cudaSuccess)
        return set_error("gemm_step_init: %s", cudaGetErrorString(e));
    return 0;
}

// Tile width / cluster size by a small cost model of the post-wait critical path (microseconds; constants measured with
// scripts/mma_rate.cu and scripts/gemm_trace.py): weight k blocks that are not resident before the wait cost a TMA load each,
// tcgen05.mma costs 61 / 74 / 136 clk at N = 64 / 128 / 256, distributed shared memory moves about 21 B/clk per SM.
static bool pick_step_config(int rows, int N, int K, int* bn, int* S) {
    float best = 1e30f;
    for (int cand_bn : {64, 128, 256}) {
        if (N % cand_bn) continue;
        for (int cand_s : {8, 4, 2, 1}) {
            StepDev d;
            size_t smem;
            if ((N / cand_bn) * cand_s > sm_count() || cand_bn / cand_s < 4 || !plan_step(rows, K, cand_bn, cand_s, &d, &smem)) continue;
            const int slice = (K / kStepBlockK + cand_s - 1) / cand_s;
            const float mma_clk = cand_bn == 64 ? 61.f : (cand_bn == 128 ? 74.f : 136.f);
            const float t = 0.3f * (float)(slice > d.w_stages ? slice - d.w_stages : 0) + slice * 4 * mma_clk / 1900.f +
                            (float)d.recv_rows * cand_bn * 4 * (cand_s - 1) / cand_s / (21.f * 1900.f) + (cand_s > 1 ? 0.4f : 0.f);
            if (t < best) {
                best = t;
                *bn = cand_bn;
                *S = cand_s;
            }
        }
    }
    return best < 1e29f;
}

// true when launch_gemm_step can take this problem (otherwise the caller uses the persistent kernel)
bool gemm_step_supported(int rows, int N, int K) {
    int bn, S;
    return rows >= 1 && rows <= 128 && N % 64 == 0 && K % 64 == 0 && K >= 256 && pick_step_config(rows, N, K, &bn, &S);
}

int launch_gemm_step(const StepGemmArgs& a, cudaStream_t stream) {
    if (!gemm_step_supported(a.rows, a.N, a.K)) return set_error("gemm_step: unsupported shape rows=%d N=%d K=%d", a.rows, a.N, a.K);
    if (a.a_row_stride % 8 || a.ldw % 8) return set_error("gemm_step: strides must be multiples of 8 halfs");
    int bn = a.block_n, S = a.cluster;
    if (bn == 0 || S == 0) pick_step_config(a.rows, a.N, a.K, &bn, &S);
    if (a.N % bn) return set_error("gemm_step: N=%d not a multiple of the tile width %d", a.N, bn);
    if (S != 1 && S != 2 && S != 4 && S != 8) return set_error("gemm_step: cluster size %d", S);
    if (bn / S < 4) return set_error("gemm_step: tile %d too narrow for %d slices", bn, S);
    if ((a.K + kStepBlockK - 1) / kStepBlockK < S) return set_error("gemm_step: K=%d too short for %d slices", a.K, S);
    if (a.out_row_stride % 4) return set_error("gemm_step: output row stride must be a multiple of 4 halfs");
    if ((a.ln_gamma == nullptr) != (a.ln_beta == nullptr)) return set_error("gemm_step: LayerNorm needs both gamma and beta");
    if (a.ln_stats && !a.ln_gamma) return set_error("gemm_step: row statistics without a LayerNorm");
    switch (bn) {
        case 64: return launch_step_bn<64>(a, S, stream);
        case 128: return launch_step_bn<128>(a, S, stream);
        case 256: return launch_step_bn<256>(a, S, stream);
    }
    return set_error("gemm_step: unsupported tile width %d", bn);
}

}  // namespace wjb
