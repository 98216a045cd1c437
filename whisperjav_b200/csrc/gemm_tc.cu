// tcgen05 + TMA persistent GEMM for sm_100a:  out[r, n] = epilogue( sum_k A[r, k] * W[n, k] )
//
// Replaces the cuBLAS/cuDNN calls the reference reaches through torch (openai-whisper
// model.py::Linear / Conv1d, called from whisperjav/modules/whisper_pro_asr.py:433).
//
//  * A is addressed through a 3-D TMA tensor map (k, row, batch) so that the k=3 convolutions
//    run as im2col-free GEMMs: with channels-last activations [batch][T+2][C] the im2col row of
//    output frame t is the contiguous span of 3*C halfs starting at padded row stride*t, i.e. a
//    tensor whose row stride (stride*C) is smaller than its row length (3*C).
//  * W is [N][K] row-major (torch Linear layout), 2-D tensor map.  Both operands are K-major,
//    SWIZZLE_128B, 64-half (128 B) k-blocks.
//  * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM
//    allocator, warps 4..11 = epilogue (TMEM -> registers -> fused bias/GELU/residual -> HBM).
//    Two TMEM accumulator buffers let the epilogue of tile i overlap the MMAs of tile i+1.
//  * Epilogue keeps the rounding points of the reference's fp16 run: half(acc + bias), then
//    GELU / residual / positional add each on the fp16-rounded value (see oracle/whisper_oracle.py).
#include "kernels.h"

namespace wjb {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;   // 64 halfs = 128 B = one SWIZZLE_128B row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiWarps = 8;
constexpr int kStgRowBytes = 80;  // 64 B of fp16 + 16 B pad: conflict-free 16-byte accesses both ways

template <int BN>
struct GemmCfg {
    static constexpr int kStageBytesA = kBlockM * kBlockK * 2;
    static constexpr int kStageBytesB = BN * kBlockK * 2;
    static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
    static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : (BN == 64 ? 8 : 10));
    static constexpr int kTmemCols = 2 * BN;
    static constexpr int kEpiWarpsActive = (BN >= 64) ? kNumEpiWarps : 4;  // BN = 32: one 32-column chunk per lane quadrant
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kNumEpiWarps * 32 * kStgRowBytes;
};

struct GemmDev {
    const __half* bias;      // [N] or null
    const __half* residual;  // indexed like out, or null
    const float* pos;        // [rows_per_batch][N] fp32, added after GELU (encoder conv2), or null
    __half* out;
    long long out_row_stride;    // elements
    long long out_batch_stride;  // elements
    int rows_per_batch, n_batch, N, K;
    int flags;
    int hs_T, hs_H;  // head-split output: row (b, t) col c -> ((b*H + c/64)*T + t)*64 + c%64
    int splits;      // split-K: work item = (tile, k slice); partial tiles meet in sk_ws, the last arriver finishes the tile
    float* sk_ws;    // [tiles][splits][128][BN] fp32
    unsigned* sk_cnt;  // [tiles], zero between launches
    unsigned long long* trace;  // debugging aid: per-launch timeline of CTA 0 (null = off)
};

// timeline slot k of this launch: SM clock and global timer
#define WJB_TRACE(k)                                                                  \
    do {                                                                              \
        if (trace_row) {                                                              \
            unsigned long long gt_;                                                   \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                   \
            trace_row[2 * (k)] = clock64();                                           \
            trace_row[2 * (k) + 1] = gt_;                                             \
        }                                                                             \
    } while (0)

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDev p) {
    using Cfg = GemmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + Cfg::kStages;
    uint64_t* tfull_bar = bars + 2 * Cfg::kStages;
    uint64_t* tempty_bar = bars + 2 * Cfg::kStages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 4);
    uint32_t* sk_flag = tmem_slot + 1;
    uint8_t* stage_base = smem + Cfg::kStages * Cfg::kStageBytes + 256;  // epilogue transpose staging, 8 warps x 32 rows x 80 B

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // PDL: the next kernel may start launching
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    __shared__ unsigned long long* trace_row_sh;
    if (p.trace && threadIdx.x == 0) {
        unsigned long long* r = nullptr;
        if (blockIdx.x == 0) {
            const unsigned long long idx = atomicAdd(p.trace, 1ull);
            r = p.trace + 32 + (idx % 512) * 32;  // ring of the last 512 launches
            {
                unsigned smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                r[30] = smid;
                r[29] = idx;
                r[28] = ((unsigned long long)p.N << 32) | (unsigned)p.K;
            }
        }
        trace_row_sh = r;
    }
    unsigned long long* trace_row = nullptr;
    if (p.trace) {
        if (threadIdx.x == 0) trace_row = trace_row_sh;
        WJB_TRACE(0);
    }
    const int m_tiles_per_batch = (p.rows_per_batch + kBlockM - 1) / kBlockM;
    const int tiles_m = m_tiles_per_batch * p.n_batch;
    const int tiles_n = (p.N + BN - 1) / BN;
    const int splits = p.splits > 1 ? p.splits : 1;
    const int num_tiles = tiles_m * tiles_n * splits;  // work items: the k slices of a tile are adjacent
    const int num_kb_total = (p.K + kBlockK - 1) / kBlockK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], Cfg::kEpiWarpsActive);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (p.trace) trace_row = trace_row_sh;  // every role's elected lane may stamp
    if (threadIdx.x == 0) WJB_TRACE(1);
    asm volatile("griddepcontrol.wait;" ::: "memory");  // PDL: inputs of the previous kernel are complete and visible
    if (threadIdx.x == 0) WJB_TRACE(2);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
                const int tile = item / splits, ks = item % splits;
                const int nt = tile % tiles_n, mt = tile / tiles_n;
                const int b = mt / m_tiles_per_batch;
                const int row0 = (mt % m_tiles_per_batch) * kBlockM;
                const int kb0 = num_kb_total * ks / splits, kb1 = num_kb_total * (ks + 1) / splits;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::kStageBytes;
                    uint8_t* sb = sa + Cfg::kStageBytesA;
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    tma_load_3d(sa, &tmA, &full_bar[stage], kb * kBlockK, row0, b);
                    tma_load_2d(sb, &tmB, &full_bar[stage], kb * kBlockK, nt * BN);
                    if (kb == kb0 && item == (int)blockIdx.x) WJB_TRACE(3);
                    if (++stage == Cfg::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(kBlockM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
                const int ks = item % splits;
                const int num_kb = num_kb_total * (ks + 1) / splits - num_kb_total * ks / splits;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (kb == 0 && item == (int)blockIdx.x) WJB_TRACE(4);
                    const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint32_t sb = sa + Cfg::kStageBytesA;
                    const uint64_t da = make_smem_desc(sa, 16, 1024, kLayoutSW128);
                    const uint64_t db = make_smem_desc(sb, 16, 1024, kLayoutSW128);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                        // advance 32 B (16 halfs) inside the 128 B swizzle row: +2 in the >>4 address field
                        umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == Cfg::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tfull_bar[acc]);
                if (item == (int)blockIdx.x) WJB_TRACE(5);
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + Cfg::kEpiWarpsActive) {
        // ===================== epilogue =====================
        const int q = warp & 3;                   // TMEM lane quadrant this warp may read
        const int ch = (warp - kEpiWarp0) >> 2;   // column half
        constexpr int kColsPerWarp = (BN >= 64) ? BN / 2 : BN;
        uint8_t* stg = stage_base + (warp - kEpiWarp0) * (32 * kStgRowBytes);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
            const int tile = item / splits, ks = item % splits;
            const int nt = tile % tiles_n, mt = tile / tiles_n;
            const int b = mt / m_tiles_per_batch;
            const int row_base = (mt % m_tiles_per_batch) * kBlockM + q * 32;  // first of this warp's 32 rows
            mbar_wait_warp(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            if (warp == kEpiWarp0 && lane == 0 && item == (int)blockIdx.x) WJB_TRACE(6);
            const uint32_t taddr0 = tmem_base + (uint32_t(q * 32) << 16) + acc * BN + ch * kColsPerWarp;

            // bias / GELU / positional / residual on one 32-column chunk of this thread's row, then the transposed store
            auto finish_chunk = [&](float (&v)[32], int c, const uint4 (&resv)[4]) {
                const int col0 = nt * BN + ch * kColsPerWarp + c;
                if (p.bias) {
                    const uint4* bp = reinterpret_cast<const uint4*>(p.bias + col0);
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        uint4 bb = __ldg(bp + j4);
                        const __half2* h2 = reinterpret_cast<const __half2*>(&bb);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float2 f = __half22float2(h2[j]);
                            v[j4 * 8 + 2 * j] += f.x;
                            v[j4 * 8 + 2 * j + 1] += f.y;
                        }
                    }
                }
                if (p.flags & GEMM_GELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(round_f16(v[j]));
                }
                // ---- transpose through smem so that global accesses are 64 B contiguous per row
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    uint4 o;
                    __half2* h2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
                    for (int j = 0; j < 4; ++j) h2[j] = __floats2half2_rn(v[j4 * 8 + 2 * j], v[j4 * 8 + 2 * j + 1]);
                    *reinterpret_cast<uint4*>(stg + lane * kStgRowBytes + j4 * 16) = o;
                }
                __syncwarp();
                const int piece = lane & 3;
                const int colp = col0 + piece * 8;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rr = (lane >> 2) + 8 * i;
                    const int grow = row_base + rr;
                    if (grow < p.rows_per_batch && colp < p.N) {
                        uint4 val = *reinterpret_cast<const uint4*>(stg + rr * kStgRowBytes + piece * 16);
                        long long off;
                        if (p.flags & GEMM_HEADSPLIT) {
                            off = ((long long)(b * p.hs_H + colp / 64) * p.hs_T + grow) * 64 + (colp % 64);
                        } else {
                            off = (long long)b * p.out_batch_stride + (long long)grow * p.out_row_stride + colp;
                        }
                        if (p.pos || p.residual) {
                            __half2* h2 = reinterpret_cast<__half2*>(&val);
                            float f[8];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float2 t = __half22float2(h2[j]);
                                f[2 * j] = t.x;
                                f[2 * j + 1] = t.y;
                            }
                            if (p.pos) {
                                const float4* pp = reinterpret_cast<const float4*>(p.pos + (long long)grow * p.N + colp);
                                const float4 a0 = __ldg(pp), a1 = __ldg(pp + 1);
                                f[0] += a0.x; f[1] += a0.y; f[2] += a0.z; f[3] += a0.w;
                                f[4] += a1.x; f[5] += a1.y; f[6] += a1.z; f[7] += a1.w;
#pragma unroll
                                for (int j = 0; j < 8; ++j) f[j] = round_f16(f[j]);
                            }
                            if (p.residual) {
                                const uint4 rv = resv[i];
                                const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float2 t = __half22float2(rh[j]);
                                    f[2 * j] += t.x;
                                    f[2 * j + 1] += t.y;
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) h2[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
                        }
                        if (colp + 8 <= p.N) {
                            *reinterpret_cast<uint4*>(p.out + off) = val;
                        } else {
                            const __half* hv = reinterpret_cast<const __half*>(&val);
                            for (int j = 0; j < 8 && colp + j < p.N; ++j) p.out[off + j] = hv[j];
                        }
                    }
                }
                __syncwarp();
            };
            // residual for the (row, 8-column piece) slots this lane owns after the transpose
            auto load_residual = [&](uint4 (&resv)[4], int c) {
                const int col0 = nt * BN + ch * kColsPerWarp + c;
#pragma unroll
                for (int i = 0; i < 4; ++i) resv[i] = make_uint4(0, 0, 0, 0);
                if (p.residual && col0 < p.N) {
                    const int colp_ = col0 + (lane & 3) * 8;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int grow_ = row_base + (lane >> 2) + 8 * i;
                        if (grow_ < p.rows_per_batch && colp_ + 8 <= p.N) {
                            const long long off_ = (long long)b * p.out_batch_stride + (long long)grow_ * p.out_row_stride + colp_;
                            resv[i] = *reinterpret_cast<const uint4*>(p.residual + off_);
                        }
                    }
                }
            };

            if (splits == 1) {
#pragma unroll 1
                for (int c = 0; c < kColsPerWarp; c += 32) {
                    uint4 resv[4];
                    load_residual(resv, c);  // issued before the TMEM read so the latency hides behind it
                    uint32_t r[32];
                    tmem_ld_32x32(taddr0 + c, r);
                    tmem_ld_wait();
                    if (nt * BN + ch * kColsPerWarp + c < p.N) {  // warp-uniform
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                        finish_chunk(v, c, resv);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            } else {
                // ---- split-K: park the raw fp32 tile, the last k slice to arrive adds the slices in order and finishes
                float* wtile = p.sk_ws + ((size_t)tile * splits) * (kBlockM * BN);
                const int trow = q * 32 + lane;
                const bool row_ok = row_base + lane < p.rows_per_batch;  // rows past the batch are never stored
#pragma unroll 1
                for (int c = 0; c < kColsPerWarp; c += 32) {
                    if (row_base >= p.rows_per_batch) break;  // warp-uniform: nothing of this quadrant is live
                    uint32_t r[32];
                    tmem_ld_32x32(taddr0 + c, r);
                    tmem_ld_wait();
                    if (!row_ok) continue;
                    float4* dst = reinterpret_cast<float4*>(wtile + (size_t)ks * (kBlockM * BN) + (size_t)trow * BN + ch * kColsPerWarp + c);
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        dst[j4] = make_float4(__uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1]), __uint_as_float(r[4 * j4 + 2]),
                                              __uint_as_float(r[4 * j4 + 3]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                __threadfence();
                asm volatile("bar.sync 1, %0;" ::"n"(Cfg::kEpiWarpsActive * 32) : "memory");
                if (warp == kEpiWarp0 && lane == 0) {
                    const unsigned old = atomicAdd(p.sk_cnt + tile, 1u);
                    const bool last = (old == (unsigned)splits - 1u);
                    if (last) p.sk_cnt[tile] = 0;  // ready for the next launch
                    *sk_flag = last ? 1u : 0u;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(Cfg::kEpiWarpsActive * 32) : "memory");
                const bool last = (*sk_flag != 0u);
                asm volatile("bar.sync 1, %0;" ::"n"(Cfg::kEpiWarpsActive * 32) : "memory");  // flag read by all before it is reused
                if (last) {
                    __threadfence();
#pragma unroll 1
                    for (int c = 0; c < kColsPerWarp; c += 32) {
                        if (nt * BN + ch * kColsPerWarp + c >= p.N || row_base >= p.rows_per_batch) continue;
                        uint4 resv[4];
                        load_residual(resv, c);
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.f;
                        for (int s_ = 0; s_ < splits && row_ok; ++s_) {
                            const float4* src = reinterpret_cast<const float4*>(wtile + (size_t)s_ * (kBlockM * BN) + (size_t)trow * BN + ch * kColsPerWarp + c);
#pragma unroll
                            for (int j4 = 0; j4 < 8; ++j4) {
                                const float4 t = __ldcg(src + j4);
                                v[4 * j4] += t.x;
                                v[4 * j4 + 1] += t.y;
                                v[4 * j4 + 2] += t.z;
                                v[4 * j4 + 3] += t.w;
                            }
                        }
                        finish_chunk(v, c, resv);
                    }
                }
            }
            if (warp == kEpiWarp0 && lane == 0 && item == (int)blockIdx.x) WJB_TRACE(7);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) WJB_TRACE(8);
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------ host side
static int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
    auto fn = get_tensor_map_encoder();
    if (!fn) return set_error("cuTensorMapEncodeTiled unavailable");
    uint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled failed (%d) rank=%d dims=%llu,%llu", (int)r, rank,
                                            (unsigned long long)dims[0], (unsigned long long)dims[1]);
    return 0;
}

template <int BN>
static int launch_bn(const GemmArgs& a, const CUtensorMap& tmA, const GemmDev& d, int num_tiles, cudaStream_t stream) {
    CUtensorMap tmB;
    uint64_t dimsB[2] = {(uint64_t)a.K, (uint64_t)a.N};
    uint64_t strB[1] = {(uint64_t)a.ldw * 2};
    uint32_t boxB[2] = {(uint32_t)kBlockK, (uint32_t)BN};
    if (int e = encode_map(&tmB, a.W, 2, dimsB, strB, boxB)) return e;
    int grid = num_tiles < sm_count() ? num_tiles : sm_count();
    cudaError_t e = launch_k(gemm_tc_kernel<BN>, dim3(grid), dim3(kGemmThreads), (size_t)GemmCfg<BN>::kSmemBytes, stream, tmA, tmB, d);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("gemm_tc launch: %s", cudaGetErrorString(e));
    return 0;
}

int gemm_init() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<256>::kSmemBytes)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128>::kSmemBytes)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<64>::kSmemBytes)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(gemm_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<32>::kSmemBytes)) != cudaSuccess)
        return set_error("gemm_init: %s", cudaGetErrorString(e));
    return 0;
}

static unsigned long long* g_trace = nullptr;
void gemm_set_trace(void* buf) { g_trace = reinterpret_cast<unsigned long long*>(buf); }

int launch_gemm(const GemmArgs& a, cudaStream_t stream) {
    if ((a.bias || a.residual || a.pos) && (a.N % 32)) return set_error("gemm: bias/residual/pos need N %% 32 == 0");
    if (a.K % 8 != 0 || a.ldw % 8 != 0) return set_error("gemm: K and ldw must be multiples of 8 (16-byte TMA rows)");
    if ((a.a_row_stride % 8) || (a.a_batch_stride % 8)) return set_error("gemm: A strides must be multiples of 8 halfs");
    CUtensorMap tmA;
    uint64_t dimsA[3] = {(uint64_t)a.K, (uint64_t)a.rows_per_batch, (uint64_t)a.n_batch};
    uint64_t strA[2] = {(uint64_t)a.a_row_stride * 2, (uint64_t)(a.n_batch > 1 ? a.a_batch_stride : a.a_row_stride * (long long)a.rows_per_batch) * 2};
    if (strA[1] == 0) strA[1] = 16;
    uint32_t boxA[3] = {(uint32_t)kBlockK, (uint32_t)kBlockM, 1};
    if (int e = encode_map(&tmA, a.A, 3, dimsA, strA, boxA)) return e;
    GemmDev d;
    d.bias = a.bias;
    d.residual = a.residual;
    d.pos = a.pos;
    d.out = a.out;
    d.out_row_stride = a.out_row_stride;
    d.out_batch_stride = a.out_batch_stride;
    d.rows_per_batch = a.rows_per_batch;
    d.n_batch = a.n_batch;
    d.N = a.N;
    d.K = a.K;
    d.flags = a.flags;
    d.hs_T = a.hs_T;
    d.hs_H = a.hs_H;
    const int tiles_m = ((a.rows_per_batch + kBlockM - 1) / kBlockM) * a.n_batch;
    int bn = a.block_n;
    if (bn == 0) {
        // enough tiles to fill the machine with the widest tile that still gives >= 1 wave
        bn = 256;
        if (tiles_m * ((a.N + 255) / 256) < sm_count()) bn = 128;
        if (tiles_m * ((a.N + 127) / 128) < sm_count()) bn = 64;
        if (tiles_m * ((a.N + 63) / 64) * 2 <= sm_count()) bn = 32;
    }
    const int tiles = tiles_m * ((a.N + bn - 1) / bn);
    d.trace = g_trace;
    d.splits = 1;
    d.sk_ws = nullptr;
    d.sk_cnt = nullptr;
    if (a.splits != 0 && a.splits != 1) {
        if (!a.splitk_ws || !a.splitk_cnt) return set_error("gemm: split-K needs a workspace");
        const int num_kb = (a.K + kBlockK - 1) / kBlockK;
        int sp = a.splits > 1 ? a.splits : sm_count() / tiles;
        if (sp > num_kb / 2) sp = num_kb / 2;  // at least two k blocks per slice
        const size_t per_split = (size_t)tiles * kBlockM * bn * sizeof(float);
        if (sp > 1 && per_split * sp > a.splitk_ws_bytes) sp = (int)(a.splitk_ws_bytes / per_split);
        if (tiles > kSplitKCounters) sp = 1;
        if (sp > 1) {
            d.splits = sp;
            d.sk_ws = a.splitk_ws;
            d.sk_cnt = a.splitk_cnt;
        }
    }
    const int items = tiles * d.splits;
    switch (bn) {
        case 256: return launch_bn<256>(a, tmA, d, items, stream);
        case 128: return launch_bn<128>(a, tmA, d, items, stream);
        case 64: return launch_bn<64>(a, tmA, d, items, stream);
        case 32: return launch_bn<32>(a, tmA, d, items, stream);
    }
    return set_error("gemm: unsupported block_n %d", bn);
}

}  // namespace wjb
