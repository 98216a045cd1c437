// Attention kernels.
//
//  * attn_encoder_kernel: non-causal flash attention for the Whisper encoder (T = 1500, d = 64),
//    tcgen05 for both matmuls.  One CTA = 128 query rows of one (batch, head); S = Q K^T lands in
//    TMEM, four softmax warps (thread == query row == TMEM lane) read it back with tcgen05.ld, write
//    fp16 P into a SWIZZLE_128B K-major smem tile, and P V runs as a second UMMA with V consumed
//    MN-major straight from the TMA tile (no transpose).  The running output is kept in registers
//    (O = O * alpha + PV).  Two CTAs are co-resident per SM (256 TMEM columns, ~97 KB smem each) so
//    one CTA's exp phase overlaps the other's MMAs.  Replaces F.scaled_dot_product_attention in
//    openai-whisper model.py::MultiHeadAttention.qkv_attention.
//  * attn_dec_self_kernel / attn_dec_cross_kernel: single-query decode attention against the HBM
//    KV cache / the per-window cross K,V; pure streaming, HBM-bound (cross: 384 KB per (b, head)).
#include <stdlib.h>

#include "kernels.h"

namespace wjb {

// ============================================================================ encoder
constexpr int kAttnThreads = 192;
constexpr int kQTile = 128, kKVTile = 128, kHeadDim = 64;
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KB
constexpr int kAttnSmem = 1024 + kTileBytes /*Q*/ + 2 * kTileBytes /*K*/ + kTileBytes /*V*/ + 2 * kTileBytes /*P*/ + 256;
constexpr int kAttnTmemCols = 256;  // S: [0,128), O: [128,192)

__global__ void __launch_bounds__(kAttnThreads, 2)
attn_encoder_kernel(const __grid_constant__ CUtensorMap tmQKV, __half* __restrict__ out, int T, int n_state) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + kTileBytes;       // 2 stages
    uint8_t* sV = sK + 2 * kTileBytes;
    uint8_t* sP = sV + kTileBytes;       // 2 k-blocks of [128 rows][64 keys]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kTileBytes);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;   // [2]
    uint64_t* k_empty = bars + 3;  // [2]
    uint64_t* v_full = bars + 5;
    uint64_t* pv_done = bars + 6;  // PV_j finished: V stage free, P tile free, O accumulator up to date
    uint64_t* s_full = bars + 7;
    uint64_t* s_free = bars + 8;   // softmax warps hold S_j in registers: TMEM S may be overwritten
    uint64_t* p_full = bars + 9;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kQTile, h = blockIdx.y, b = blockIdx.z;
    const int nkv = (T + kKVTile - 1) / kKVTile;

    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
        }
        mbar_init(v_full, 1);
        mbar_init(pv_done, 1);
        mbar_init(s_full, 1);
        mbar_init(s_free, 128);
        mbar_init(p_full, 128);
        fence_barrier_init();
    }
    if (warp == 0) {
        if (lane == 0) tma_prefetch_desc(&tmQKV);
        tmem_alloc<kAttnTmemCols>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, kTileBytes);
            tma_load_3d(sQ, &tmQKV, q_full, h * kHeadDim, q0, b);
            for (int j = 0; j < nkv; ++j) {
                const int st = j & 1;
                mbar_wait(&k_empty[st], ((j >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&k_full[st], kTileBytes);
                tma_load_3d(sK + st * kTileBytes, &tmQKV, &k_full[st], n_state + h * kHeadDim, j * kKVTile, b);
                mbar_wait(pv_done, (j & 1) ^ 1);  // PV_{j-1} done -> V stage free
                mbar_arrive_expect_tx(v_full, kTileBytes);
                tma_load_3d(sV, &tmQKV, v_full, 2 * n_state + h * kHeadDim, j * kKVTile, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0);  // S = Q K^T, both K-major
            constexpr uint32_t idesc_o = make_idesc_f16(128, 64, 0, 1);   // O += P V, V is MN-major
            mbar_wait(q_full, 0);
            const uint64_t dq = make_smem_desc(smem_u32(sQ), 16, 1024, kLayoutSW128);
            // software pipeline: S_j is issued one block ahead of PV_{j-1}
            for (int j = 0; j <= nkv; ++j) {
                if (j < nkv) {
                    const int st = j & 1;
                    mbar_wait(&k_full[st], (j >> 1) & 1);
                    if (j > 0) mbar_wait(s_free, (j - 1) & 1);
                    tc_fence_after();
                    const uint64_t dk = make_smem_desc(smem_u32(sK + st * kTileBytes), 16, 1024, kLayoutSW128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(tmem_S, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                    umma_commit(s_full);
                    umma_commit(&k_empty[st]);
                }
                if (j > 0) {
                    const int jj = j - 1;
                    mbar_wait(p_full, jj & 1);
                    mbar_wait(v_full, jj & 1);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        // A: P k-block (k / 4), 32 B per 16 keys inside the 128 B row.  B: V rows 16*k.. (128 B per key row).
                        const uint64_t dp = make_smem_desc(smem_u32(sP + (k >> 2) * kTileBytes) + (k & 3) * 32, 16, 1024, kLayoutSW128);
                        const uint64_t dv = make_smem_desc(smem_u32(sV) + k * 16 * 128, 1024, 1024, kLayoutSW128);
                        umma_f16(tmem_O, dp, dv, idesc_o, (jj | k) != 0);
                    }
                    umma_commit(pv_done);
                }
            }
        }
    } else {
        // ===================== softmax warps: thread == query row =====================
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const uint32_t lane_off = uint32_t(quad * 32) << 16;
        const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
        float m_ref = -INFINITY, l = 0.f;                // m_ref: the max the accumulated O, l and P are scaled to

        for (int j = 0; j < nkv; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            const int kvalid = T - j * kKVTile;  // keys >= kvalid are padding (only possible in the last block)
            uint32_t pk[2][32];                   // fp16x2 P of the two 64-key halves
            bool pv_waited = (j == 0);
            // The block is consumed in two 64-key halves so that only 64 scores are live at a time.  The reference max
            // is lazy: it only moves when a half's max exceeds it by more than 2^8; then O, l and any P already packed
            // for this block are rescaled (rare after the first block).
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t s[64];
                tmem_ld_32x32(tmem_S + lane_off + hf * 64, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
                tmem_ld_32x32(tmem_S + lane_off + hf * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
                tmem_ld_wait();
                if (hf == 1) {
                    tc_fence_before();
                    mbar_arrive(s_free);  // S_j fully read: the tensor core may overwrite it with S_{j+1}
                }
                if (kvalid < kKVTile) {
#pragma unroll
                    for (int i = 0; i < 64; ++i)
                        if (hf * 64 + i >= kvalid) s[i] = 0xff800000u;  // -inf: exp2 -> 0
                }
                // four independent maxima (a 64-deep dependent FMNMX chain costs ~250 cycles per half with two softmax warps per scheduler)
                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
                for (int i = 0; i < 64; i += 4) {
                    mx0 = fmaxf(mx0, __uint_as_float(s[i]));
                    mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
                    mx2 = fmaxf(mx2, __uint_as_float(s[i + 2]));
                    mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
                }
                const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                if (j == 0 && hf == 0) {
                    m_ref = mx;
                } else {
                    const bool need = (mx - m_ref) * sl2 > 8.0f;
                    if (__any_sync(0xffffffffu, need)) {
                        if (!pv_waited) {
                            mbar_wait(pv_done, (j - 1) & 1);
                            tc_fence_after();
                            pv_waited = true;
                        }
                        const float alpha = need ? ex2_approx((m_ref - mx) * sl2) : 1.0f;
                        if (need) m_ref = mx;
                        l *= alpha;
                        if (j > 0) {
#pragma unroll 1
                            for (int c = 0; c < 4; ++c) {
                                uint32_t o[16];
                                tmem_ld_32x16(tmem_O + lane_off + c * 16, o);
                                tmem_ld_wait();
#pragma unroll
                                for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                                tmem_st_32x16(tmem_O + lane_off + c * 16, o);
                            }
                            tmem_st_wait();
                        }
                        if (hf == 1) {
                            const __half2 a2 = __float2half2_rn(alpha);
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                __half2 v = *reinterpret_cast<__half2*>(&pk[0][i]);
                                v = __hmul2(v, a2);
                                pk[0][i] = *reinterpret_cast<uint32_t*>(&v);
                            }
                        }
                    }
                }
                const float mb = m_ref * sl2;
                // two accumulation chains, in this order: the order of these fp32 adds is part of the encoder's bit pattern, which the
                // model-level parity thresholds were measured on (four chains were tried: not faster, and every downstream tie re-rolled)
                float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float p0 = ex2_approx(fmaf(__uint_as_float(s[2 * i]), sl2, -mb));
                    const float p1 = ex2_approx(fmaf(__uint_as_float(s[2 * i + 1]), sl2, -mb));
                    ls0 += p0;
                    ls1 += p1;
                    __half2 hp = __floats2half2_rn(p0, p1);
                    pk[hf][i] = *reinterpret_cast<uint32_t*>(&hp);
                }
                l += ls0 + ls1;
            }
            if (!pv_waited) mbar_wait(pv_done, (j - 1) & 1);  // P tile is free once PV_{j-1} has read it
            // half hf = k-block hf of the P tile; 16-byte chunk q of this row goes to slot q ^ (row & 7) (SWIZZLE_128B)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint8_t* prow = sP + hf * kTileBytes + row * 128;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int chunk = q ^ (row & 7);
                    *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[hf][q * 4], pk[hf][q * 4 + 1], pk[hf][q * 4 + 2], pk[hf][q * 4 + 3]);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(p_full);
        }
        mbar_wait(pv_done, (nkv - 1) & 1);
        tc_fence_after();
        const float inv_l = 1.0f / l;
        const bool row_ok = q0 + row < T;
        __half* orow = out + ((long long)b * T + q0 + row) * n_state + h * kHeadDim;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_O + lane_off + c * 32, r);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint4 o;
                    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int idx = q4 * 8 + 2 * i;
                        oh[i] = __floats2half2_rn(__uint_as_float(r[idx]) * inv_l, __uint_as_float(r[idx + 1]) * inv_l);
                    }
                    *reinterpret_cast<uint4*>(orow + c * 32 + q4 * 8) = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<kAttnTmemCols>(tmem_base);
    }
}

int attn_init() {
    cudaError_t e = cudaFuncSetAttribute(attn_encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e != cudaSuccess) return set_error("attn attr: %s", cudaGetErrorString(e));
    return 0;
}

int launch_attn_encoder(const __half* qkv, __half* out, int B, int T, int H, cudaStream_t s) {
    const int n = H * kHeadDim;
    auto fn = get_tensor_map_encoder();
    if (!fn) return set_error("cuTensorMapEncodeTiled unavailable");
    CUtensorMap tm;
    uint64_t dims[3] = {(uint64_t)3 * n, (uint64_t)T, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)3 * n * 2, (uint64_t)T * 3 * n * 2};
    uint32_t box[3] = {64, 128, 1};
    uint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(qkv), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("attn tensor map encode failed (%d)", (int)r);
    dim3 grid((T + kQTile - 1) / kQTile, H, B);
    attn_encoder_kernel<<<grid, kAttnThreads, kAttnSmem, s>>>(tm, out, T, n);
    WJB_CHECK_LAUNCH("attn_encoder");
    return 0;
}

// ============================================================================ decoder self-attention
// One CTA (128 threads) per (b, head).  kv_cache [B][2H][n_ctx][64]: K heads 0..H-1, V heads H..2H-1.  Appends this
// position's k, v, then single-query attention over positions 0..pos: 8 lanes x 16 B per cached row, 16 rows per pass and
// 4 passes in flight per thread, so that at position 200+ the kernel is bound by cache bandwidth and not by one warp's
// chain of dependent row loads (the first version, one warp per head, took 40 us per layer there).
constexpr int kSelfThreads = 128;
__global__ void __launch_bounds__(kSelfThreads) attn_dec_self_kernel(const __half* __restrict__ qkv, __half* __restrict__ kv_cache,
                                                                      __half* __restrict__ out, const int* __restrict__ step_ptr,
                                                                      const unsigned char* __restrict__ done, int H, int n_ctx,
                                                                      const short* __restrict__ anc, long long anc_parity_stride) {
    pdl_prologue();
    const int h = blockIdx.x, b = blockIdx.y;
    if (done && done[b]) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pos = *step_ptr;
    const int T = pos + 1;
    const int n = H * 64;
    __shared__ float sc[448];
    // beam search: position t of this row's history lives in the cache of physical row src[t] (its ancestor at that step)
    __shared__ short src[448];
    if (anc) {
        const short* arow = anc + (pos & 1) * anc_parity_stride + (long long)b * n_ctx;
        for (int t = tid; t < pos; t += kSelfThreads) src[t] = arow[t];
        if (tid == 0) src[pos] = (short)b;
    }
    __shared__ float red[8];
    __shared__ float osum[4][64];
    const __half* row = qkv + (long long)b * 3 * n;
    uint4* K = reinterpret_cast<uint4*>(kv_cache + ((long long)(b * 2 * H + h) * n_ctx) * 64);
    uint4* V = reinterpret_cast<uint4*>(kv_cache + ((long long)(b * 2 * H + H + h) * n_ctx) * 64);
    const long long row_pitch = (long long)2 * H * n_ctx * 8;  // uint4 between the caches of consecutive rows
    // append this token's k, v (8 x 16 B each)
    if (tid < 8) K[(long long)pos * 8 + tid] = reinterpret_cast<const uint4*>(row + n + h * 64)[tid];
    if (tid >= 8 && tid < 16) V[(long long)pos * 8 + tid - 8] = reinterpret_cast<const uint4*>(row + 2 * n + h * 64)[tid - 8];
    const int chunk = tid & 7;  // which 16-byte (8 dims) slice of the 64-dim row
    const int slot = tid >> 3;  // row slot 0..15 within a pass
    float qf[8];
    {
        uint4 u = reinterpret_cast<const uint4*>(row + h * 64)[chunk];
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = __half22float2(h2[j]);
            qf[2 * j] = f.x;
            qf[2 * j + 1] = f.y;
        }
    }
    __syncthreads();  // the appended row is visible to the whole CTA
    // ---- scores
    for (int t0 = 0; t0 < T; t0 += 64) {
        uint4 u[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int t = t0 + r * 16 + slot;
            u[r] = (t < T) ? (anc ? K[((long long)src[t] - b) * row_pitch + (long long)t * 8 + chunk] : K[(long long)t * 8 + chunk])
                           : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const __half2* h2 = reinterpret_cast<const __half2*>(&u[r]);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f = __half22float2(h2[j]);
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
    __syncthreads();
    // ---- softmax over T scores
    float mx = -INFINITY;
    for (int t = tid; t < T; t += kSelfThreads) mx = fmaxf(mx, sc[t]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    float sum = 0.f;
    for (int t = tid; t < T; t += kSelfThreads) {
        const float e = __expf(sc[t] - mx);
        sc[t] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[4 + warp] = sum;
    __syncthreads();
    const float inv = 1.0f / (red[4] + red[5] + red[6] + red[7]);
    // ---- out = sum_t p[t] V[t], p rounded to fp16 as the reference's softmax output is
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int t0 = 0; t0 < T; t0 += 64) {
        uint4 u[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int t = t0 + r * 16 + slot;
            u[r] = (t < T) ? (anc ? V[((long long)src[t] - b) * row_pitch + (long long)t * 8 + chunk] : V[(long long)t * 8 + chunk])
                           : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int t = t0 + r * 16 + slot;
            const float w = (t < T) ? round_f16(sc[t] * inv) : 0.f;
            const __half2* h2 = reinterpret_cast<const __half2*>(&u[r]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f = __half22float2(h2[j]);
                acc[2 * j] = fmaf(w, f.x, acc[2 * j]);
                acc[2 * j + 1] = fmaf(w, f.y, acc[2 * j + 1]);
            }
        }
    }
    // reduce over the 4 row slots inside a warp (lane bits 3,4), then over the 4 warps
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
    }
    if (lane < 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) osum[warp][lane * 8 + j] = acc[j];
    }
    __syncthreads();
    if (tid < 64) {
        const float v = osum[0][tid] + osum[1][tid] + osum[2][tid] + osum[3][tid];
        out[(long long)b * n + h * 64 + tid] = __float2half_rn(v);
    }
}

int launch_attn_dec_self(const __half* qkv, __half* kv_cache, __half* out, const int* step, const unsigned char* done, int B, int H,
                         int n_ctx, cudaStream_t s, const short* anc, long long anc_parity_stride) {
    if (n_ctx > 448) return set_error("attn_dec_self: n_ctx %d > 448", n_ctx);
    dim3 grid(H, B);
    launch_k(attn_dec_self_kernel, grid, dim3(kSelfThreads), 0, s, qkv, kv_cache, out, step, done, H, n_ctx, anc, anc_parity_stride);
    WJB_CHECK_LAUNCH("attn_dec_self");
    return 0;
}

// ============================================================================ decoder cross-attention
// One CTA (128 threads) per (b, head) (or per (window, head) with NQ beams); K and V are [T][64] fp16, streamed once each.
constexpr int kCrossThreads = 128;
constexpr int kCrossMaxT = 1536;

// K and V stream through a 4-stage smem ring filled by cp.async.bulk (one elected thread, mbarrier
// completion), so the bytes in flight are not bounded by the LSU's outstanding-request capacity; V chunks are already in
// flight while the softmax runs.
constexpr int kCbStages = 4, kCbKeys = 128, kCbStageBytes = kCbKeys * 128;
constexpr int kCbMaxQ = 4;  // queries (beams of one window) that may share one pass over the window's K/V
// threads per CTA: 128 for one query; the beam-search instances (NQ >= 2) do NQ times the arithmetic per streamed byte and were
// latency-bound with 2 CTAs x 4 warps per SM (140 us vs 80 us per launch at NQ = 2, profiles/r2_launches.md): they run 8 warps
constexpr int cb_threads(int nq) { return nq == 1 ? 128 : 256; }
// ring depth: 4 stages for one query; 3 for the beam instances so that three 256-thread CTAs still fit one SM (62.6 KB at NQ = 2)
constexpr int cb_stages(int nq) { return nq == 1 ? kCbStages : 3; }
constexpr int cb_smem_bytes(int nq) { return cb_stages(nq) * kCbStageBytes + nq * kCrossMaxT * 4 + 64 * 4 + (cb_threads(nq) / 32) * 64 * 4 + 64; }
constexpr int kCbSmem = cb_smem_bytes(1);

// NQ = 1: one CTA per (row, head); with beam search either the row reads the K/V of window row / kv_div (any beam size), or
// NQ = beam_size queries of one window share the CTA and the window's K/V stream through shared memory once (blockIdx.y = window).
template <int NQ>
__global__ void __launch_bounds__(cb_threads(NQ)) attn_dec_cross_bulk_kernel(const __half* __restrict__ q, const __half* __restrict__ kv,
                                                                             __half* __restrict__ out,
                                                                             const unsigned char* __restrict__ done, int H, int T, int kv_div,
                                                                             const CrossCapture cap, const CrossTuning tune) {
    // K/V are constants of the decode run and `done` was written a whole step ago: only q depends on the previous kernel.  With
    // tune.early_kv the ring is primed BEFORE the programmatic-dependent-launch wait, so the stream is in flight while the query
    // projection in front of this kernel drains.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (!tune.early_kv) asm volatile("griddepcontrol.wait;" ::: "memory");
    const int h = blockIdx.x, b = blockIdx.y;
    if (threadIdx.x == 0) l2_prefetch_share(tune.pf_ptr, tune.pf_bytes, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y);
    const int row0 = b * NQ;  // first query row of this CTA
    if (done && done[row0]) return;  // the beams of a window finish together
    extern __shared__ __align__(128) uint8_t cb_smem[];
    constexpr int ST = cb_stages(NQ);
    uint8_t* ring = cb_smem;
    float* sc = reinterpret_cast<float*>(cb_smem + ST * kCbStageBytes);  // [NQ][kCrossMaxT]
    float* red = sc + NQ * kCrossMaxT;
    constexpr int TH = cb_threads(NQ), W = TH / 32, kSlots = TH / 8;  // key rows in flight per pass: one per 8 lanes
    float* osum = red + 64;                      // [W][64]
    uint64_t* full = reinterpret_cast<uint64_t*>(osum + W * 64);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = H * 64;
    const int bk = (NQ > 1) ? b : b / kv_div;  // window whose K/V this CTA reads
    const uint8_t* Kg = reinterpret_cast<const uint8_t*>(kv + ((long long)(bk * 2 * H + h) * T) * 64);
    const uint8_t* Vg = reinterpret_cast<const uint8_t*>(kv + ((long long)(bk * 2 * H + H + h) * T) * 64);
    const int nck = (T + kCbKeys - 1) / kCbKeys;  // chunks per matrix
    const int total = 2 * nck;
    const unsigned long long kv_policy = tune.evict_first ? l2_policy_evict_first() : 0ull;
    auto chunk_src = [&](int c) { return (c < nck ? Kg : Vg) + (size_t)(c % nck) * kCbStageBytes; };
    auto chunk_keys = [&](int c) { return min(kCbKeys, T - (c % nck) * kCbKeys); };
    auto issue = [&](int c) {
        const int st = c % ST;
        const uint32_t bytes = chunk_keys(c) * 128;
        mbar_arrive_expect_tx(&full[st], bytes);
        if (tune.evict_first) {  // the K/V stream is read once per step: do not let it push the layer's weights out of L2
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                             smem_u32(ring + st * kCbStageBytes)),
                         "l"(chunk_src(c)), "r"(bytes), "r"(smem_u32(&full[st])), "l"(kv_policy)
                         : "memory");
        } else {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(ring + st * kCbStageBytes)),
                         "l"(chunk_src(c)), "r"(bytes), "r"(smem_u32(&full[st]))
                         : "memory");
        }
    };
    if (tid == 0) {
        for (int i = 0; i < ST; ++i) mbar_init(&full[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0)
        for (int c = 0; c < ST && c < total; ++c) issue(c);
    if (tune.early_kv) asm volatile("griddepcontrol.wait;" ::: "memory");  // q (and `out`) belong to the chain

    const int chunk16 = tid & 7, slot = tid >> 3;  // kSlots key slots x 8 sixteen-byte pieces
    float qf[NQ][8];
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        const uint4 u = reinterpret_cast<const uint4*>(q + (long long)(row0 + i) * n + h * 64)[chunk16];
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h2[j]);
            qf[i][2 * j] = f.x;
            qf[i][2 * j + 1] = f.y;
        }
    }
    float acc[NQ][8];
#pragma unroll
    for (int i = 0; i < NQ; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float inv[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) inv[i] = 0.f;
    for (int c = 0; c < total; ++c) {
        const int st = c % ST;
        mbar_wait(&full[st], (c / ST) & 1);
        const uint8_t* base = ring + st * kCbStageBytes;
        const int keys = chunk_keys(c);
        const int key0 = (c % nck) * kCbKeys;
        if (c < nck) {
#pragma unroll
            for (int it = 0; it < kCbKeys / kSlots; ++it) {
                const int k = it * kSlots + slot;
                const uint4 u = *reinterpret_cast<const uint4*>(base + k * 128 + chunk16 * 16);
                const __half2* h2 = reinterpret_cast<const __half2*>(&u);
                float kf[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(h2[j]);
                    kf[2 * j] = f.x;
                    kf[2 * j + 1] = f.y;
                }
#pragma unroll
                for (int i = 0; i < NQ; ++i) {
                    float s_ = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) s_ = fmaf(qf[i][j], kf[j], s_);
                    s_ += __shfl_xor_sync(0xffffffffu, s_, 1);
                    s_ += __shfl_xor_sync(0xffffffffu, s_, 2);
                    s_ += __shfl_xor_sync(0xffffffffu, s_, 4);
                    if (chunk16 == 0 && k < keys) sc[i * kCrossMaxT + key0 + k] = s_ * 0.125f;
                }
            }
        } else {
#pragma unroll
            for (int it = 0; it < kCbKeys / kSlots; ++it) {
                const int k = it * kSlots + slot;
                const uint4 u = (k < keys) ? *reinterpret_cast<const uint4*>(base + k * 128 + chunk16 * 16) : make_uint4(0, 0, 0, 0);
                const __half2* h2 = reinterpret_cast<const __half2*>(&u);
                float vf[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(h2[j]);
                    vf[2 * j] = f.x;
                    vf[2 * j + 1] = f.y;
                }
#pragma unroll
                for (int i = 0; i < NQ; ++i) {
                    const float w = (k < keys) ? round_f16(sc[i * kCrossMaxT + key0 + k] * inv[i]) : 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(w, vf[j], acc[i][j]);
                }
            }
        }
        __syncthreads();  // stage consumed by everyone (and, after the last K chunk, all scores are in smem)
        if (tid == 0 && c + ST < total) issue(c + ST);
        if (c == nck - 1) {
            if (NQ == 1 && cap.base) {
                // word-timestamp alignment pass: the scaled scores q.k / sqrt(d) of this (row, head, position), fp16 as the
                // reference's fp16 attention matmul returns them (timing.py collects them through forward hooks)
                __half* dst = cap.base + (long long)b * cap.b_stride + (long long)h * cap.head_stride + (long long)(*cap.step) * T;
                for (int t = tid; t < T; t += TH) dst[t] = __float2half_rn(sc[t]);
            }
            // softmax over the T scores of every query (V chunks are already streaming into the ring)
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                float* sci = sc + i * kCrossMaxT;
                float mx = -INFINITY;
                for (int t = tid; t < T; t += TH) mx = fmaxf(mx, sci[t]);
                mx = warp_max(mx);
                if (i > 0) __syncthreads();  // red[] of the previous query has been read
                if (lane == 0) red[warp] = mx;
                __syncthreads();
                mx = red[0];
#pragma unroll
                for (int w_ = 1; w_ < W; ++w_) mx = fmaxf(mx, red[w_]);
                float sum = 0.f;
                for (int t = tid; t < T; t += TH) {
                    const float e = __expf(sci[t] - mx);
                    sci[t] = e;
                    sum += e;
                }
                sum = warp_sum(sum);
                if (lane == 0) red[W + warp] = sum;
                __syncthreads();
                float tot = red[W];
#pragma unroll
                for (int w_ = 1; w_ < W; ++w_) tot += red[W + w_];
                inv[i] = 1.0f / tot;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], 8);
            acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], 16);
        }
        if (i > 0) __syncthreads();  // osum of the previous query has been read
        if (lane < 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) osum[warp * 64 + lane * 8 + j] = acc[i][j];
        }
        __syncthreads();
        if (tid < 64) {
            float v = osum[tid];
#pragma unroll
            for (int w_ = 1; w_ < W; ++w_) v += osum[w_ * 64 + tid];
            out[(long long)(row0 + i) * n + h * 64 + tid] = __float2half_rn(v);
        }
    }
}

template <int NQ>
static int launch_cross_bulk(const __half* q, const __half* kv, __half* out, const unsigned char* done, int B, int H, int T, int kv_div,
                             cudaStream_t s, const CrossCapture& cap, const CrossTuning& tune) {
    dim3 grid(H, B / NQ);
    cudaError_t e = launch_k(attn_dec_cross_bulk_kernel<NQ>, grid, dim3(cb_threads(NQ)), (size_t)cb_smem_bytes(NQ), s, q, kv, out, done, H, T, kv_div, cap,
                             tune);
    if (e != cudaSuccess) return set_error("attn_dec_cross_bulk launch: %s", cudaGetErrorString(e));
    return 0;
}

// set the dynamic shared memory attribute of every instance up front (outside any stream capture)
int attn_cross_init() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(attn_dec_cross_bulk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cb_smem_bytes(1))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(attn_dec_cross_bulk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cb_smem_bytes(2))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(attn_dec_cross_bulk_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, cb_smem_bytes(3))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(attn_dec_cross_bulk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, cb_smem_bytes(4))) != cudaSuccess)
        return set_error("cross attr: %s", cudaGetErrorString(e));
    return 0;
}

int launch_attn_dec_cross(const __half* q, const __half* kv, __half* out, const unsigned char* done, int B, int H, int T,
                          cudaStream_t s, int kv_div, const CrossCapture* capture, const CrossTuning* tuning) {
    const CrossCapture cap = capture ? *capture : CrossCapture{};
    const CrossTuning tune = tuning ? *tuning : CrossTuning{};
    if (cap.base && kv_div != 1) return set_error("attn_dec_cross: score capture needs one query per window");
    if (kv_div < 1) kv_div = 1;
    if (T > kCrossMaxT) return set_error("attn_dec_cross: T %d > %d", T, kCrossMaxT);
    // beam search: the beams of a window share one pass over its K/V when they fit one CTA
    if (kv_div > 1 && kv_div <= kCbMaxQ && B % kv_div == 0) {
        switch (kv_div) {
            case 2: return launch_cross_bulk<2>(q, kv, out, done, B, H, T, kv_div, s, cap, tune);
            case 3: return launch_cross_bulk<3>(q, kv, out, done, B, H, T, kv_div, s, cap, tune);
            case 4: return launch_cross_bulk<4>(q, kv, out, done, B, H, T, kv_div, s, cap, tune);
        }
    }
    return launch_cross_bulk<1>(q, kv, out, done, B, H, T, kv_div, s, cap, tune);
}

}  // namespace wjb
