// Word-level timestamps on the device: the arithmetic of openai-whisper timing.py::find_alignment after the teacher-forced
// decoder pass has left the cross-attention scores of the alignment heads in HBM (attn_dec_cross_bulk_kernel's capture):
//
//   weights = softmax(QK[:, :, : num_frames // 2], dim=-1)            (per head, per token row, over the content frames)
//   weights = (weights - mean_over_tokens) / std_over_tokens            (per head, per frame column; population std)
//   weights = median_filter(weights, 7)                                 (along frames, reflect padding)
//   matrix  = weights.mean(over heads)[len(sot_sequence) : -1]
//   text_indices, time_indices = dtw(-matrix)                           (upstream: a Triton kernel on CUDA, numba on CPU)
//
// Three kernels, all HBM-streaming (the captured scores are read exactly twice):
//   align_rowstat_kernel   one warp per (window, head, token row): max and 1 / sum exp over the content frames
//   align_matrix_kernel    one CTA per (window, strip of 64 frames): loops over the heads; tile [tokens x (64 + 6)] in shared
//                          memory -> column statistics -> normalise -> 7-tap median along the row -> running mean over heads
//   align_dtw_kernel       one CTA per window: anti-diagonal wavefront over the [rows x frames] cost matrix (one thread per
//                          row, three rolling diagonals in shared memory, trace bytes in HBM), then the backtrace by one
//                          thread, which leaves for every row the first frame of its path segment (= timing.py's jump_times)
#include "kernels.h"

namespace wjb {

constexpr int kAlignStrip = 64;
constexpr int kAlignHalo = 3;  // median width 7
constexpr int kAlignTileW = kAlignStrip + 2 * kAlignHalo;
constexpr int kAlignMaxRows = 448;

// qk fp16 [B][n_sel][n_steps][T]; stat fp32 [B][n_sel][n_steps][2]
__global__ void align_rowstat_kernel(const __half* __restrict__ qk, float* __restrict__ stat, const int* __restrict__ n_tok,
                                     const int* __restrict__ n_frames2, int n_sel, int n_steps, int T) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int L = n_tok[b], nf = n_frames2[b];
    if (warp >= n_sel * L) return;
    const int sel = warp / L, i = warp % L;
    const __half* row = qk + (((long long)b * n_sel + sel) * n_steps + i) * T;
    float mx = -INFINITY;
    for (int t = lane; t < nf; t += 32) mx = fmaxf(mx, __half2float(row[t]));
    mx = warp_max(mx);
    float sum = 0.f;
    for (int t = lane; t < nf; t += 32) sum += expf(__half2float(row[t]) - mx);
    sum = warp_sum(sum);
    if (lane == 0) {
        float* o = stat + ((((long long)b * n_sel + sel) * n_steps + i) << 1);
        o[0] = mx;
        o[1] = 1.0f / sum;
    }
}

__device__ __forceinline__ void cswap(float& a, float& b) {
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo;
    b = hi;
}
// median of 7 (selection network: only element 3 of the sorted order is needed)
__device__ __forceinline__ float median7(float a0, float a1, float a2, float a3, float a4, float a5, float a6) {
    cswap(a0, a5); cswap(a0, a3); cswap(a1, a6); cswap(a2, a4); cswap(a0, a1); cswap(a3, a5); cswap(a2, a6);
    cswap(a2, a3); cswap(a3, a6); cswap(a4, a5); cswap(a1, a4); cswap(a1, a3); cswap(a3, a4);
    return a3;
}

// matrix fp32 [B][n_steps][T] (rows < n_tok[b], columns < n_frames2[b] written)
__global__ void __launch_bounds__(256) align_matrix_kernel(const __half* __restrict__ qk, const float* __restrict__ stat,
                                                            float* __restrict__ matrix, const int* __restrict__ n_tok,
                                                            const int* __restrict__ n_frames2, int n_sel, int n_steps, int T, int medfilt) {
    extern __shared__ float al_smem[];
    const int b = blockIdx.y;
    const int L = n_tok[b], nf = n_frames2[b];
    const int t0 = blockIdx.x * kAlignStrip;
    if (t0 >= nf || L <= 0) return;
    float* tile = al_smem;                       // [L][kAlignTileW]
    float* acc = tile + (size_t)L * kAlignTileW;  // [L][kAlignStrip]
    float* cmean = acc + (size_t)L * kAlignStrip; // [kAlignTileW]
    float* cinv = cmean + kAlignTileW;
    const int tid = threadIdx.x;
    for (int k = tid; k < L * kAlignStrip; k += blockDim.x) acc[k] = 0.f;
    const bool filter = medfilt == 7 && nf > kAlignHalo;  // timing.py: inputs no longer than the pad width pass through unfiltered
    for (int sel = 0; sel < n_sel; ++sel) {
        const __half* base = qk + (((long long)b * n_sel + sel) * n_steps) * T;
        const float* st = stat + ((((long long)b * n_sel + sel) * n_steps) << 1);
        __syncthreads();
        // tile column c <-> frame t0 - 3 + c, reflected at 0 and nf - 1 (F.pad(mode="reflect"))
        for (int k = tid; k < L * kAlignTileW; k += blockDim.x) {
            const int i = k / kAlignTileW, c = k % kAlignTileW;
            int t = t0 - kAlignHalo + c;
            if (t < 0) t = -t;
            if (t >= nf) t = 2 * (nf - 1) - t;
            float v = 0.f;
            if (t >= 0 && t < nf) v = expf(__half2float(base[(long long)i * T + t]) - st[2 * i]) * st[2 * i + 1];
            tile[k] = v;
        }
        __syncthreads();
        // per-column mean / population std over the token rows
        if (tid < kAlignTileW) {
            float s = 0.f;
            for (int i = 0; i < L; ++i) s += tile[i * kAlignTileW + tid];
            const float m = s / L;
            float v = 0.f;
            for (int i = 0; i < L; ++i) {
                const float d = tile[i * kAlignTileW + tid] - m;
                v = fmaf(d, d, v);
            }
            cmean[tid] = m;
            cinv[tid] = 1.0f / sqrtf(v / L);
        }
        __syncthreads();
        for (int k = tid; k < L * kAlignTileW; k += blockDim.x) {
            const int c = k % kAlignTileW;
            tile[k] = (tile[k] - cmean[c]) * cinv[c];
        }
        __syncthreads();
        for (int k = tid; k < L * kAlignStrip; k += blockDim.x) {
            const int i = k / kAlignStrip, c = k % kAlignStrip;
            if (t0 + c >= nf) continue;
            const float* r = tile + i * kAlignTileW + c;  // r[3] is the centre
            acc[k] += filter ? median7(r[0], r[1], r[2], r[3], r[4], r[5], r[6]) : r[3];
        }
    }
    __syncthreads();
    const float inv = 1.0f / n_sel;
    for (int k = tid; k < L * kAlignStrip; k += blockDim.x) {
        const int i = k / kAlignStrip, c = k % kAlignStrip;
        if (t0 + c < nf) matrix[((long long)b * n_steps + i) * T + t0 + c] = acc[k] * inv;
    }
}

// One CTA per window.  Rows row_begin[b] .. row_begin[b] + n_rows[b] - 1 of -matrix are the DTW cost matrix x [N][M], M = n_frames2[b].
// trace uint8 [B][kAlignMaxRows + 1][T + 1] workspace.  jump int32 [B][n_steps]: jump[i] = first frame of row i's path segment.
__global__ void __launch_bounds__(kAlignMaxRows) align_dtw_kernel(const float* __restrict__ matrix, unsigned char* __restrict__ trace,
                                                                   int* __restrict__ jump, const int* __restrict__ row_begin,
                                                                   const int* __restrict__ n_rows, const int* __restrict__ n_frames2,
                                                                   int n_steps, int T) {
    const int b = blockIdx.x;
    const int N = n_rows[b], M = n_frames2[b], r0 = row_begin[b];
    if (N <= 0 || M <= 0) return;
    __shared__ float diag[3][kAlignMaxRows + 1];  // cost on anti-diagonals d-2, d-1, d (index = i)
    const int tid = threadIdx.x;  // row i = tid + 1
    const long long tw = T + 1;
    unsigned char* tr = trace + (long long)b * (kAlignMaxRows + 1) * tw;
    const float* x = matrix + ((long long)b * n_steps + r0) * T;
    // d = i + j, cell (i, j), 1 <= i <= N, 1 <= j <= M.  cost[0][0] = 0, cost[0][j > 0] = cost[i > 0][0] = inf.
    for (int k = tid; k <= N; k += blockDim.x) {
        diag[0][k] = INFINITY;
        diag[1][k] = INFINITY;
        diag[2][k] = INFINITY;
    }
    __syncthreads();
    if (tid == 0) diag[0][0] = 0.f;  // diagonal d = 0 holds cost[0][0]
    __syncthreads();
    // buffers: cur = d % 3, prev1 = (d - 1) % 3, prev2 = (d - 2) % 3, indexed by i; cost[i][j] lives at diag[(i + j) % 3][i]
    for (int d = 2; d <= N + M; ++d) {
        const int i = tid + 1, j = d - i;
        float* cur = diag[d % 3];
        const float* p1 = diag[(d + 2) % 3];  // d - 1
        const float* p2 = diag[(d + 1) % 3];  // d - 2
        if (i <= N && j >= 1 && j <= M) {
            const float c0 = p2[i - 1];                                   // cost[i-1][j-1]
            const float c1 = (j == 0) ? INFINITY : p1[i - 1];             // cost[i-1][j]
            const float c2 = (j - 1 == 0) ? INFINITY : p1[i];             // cost[i][j-1]  (column 0 is inf for i > 0)
            float c;
            unsigned char t;
            if (c0 < c1 && c0 < c2) {
                c = c0;
                t = 0;
            } else if (c1 < c0 && c1 < c2) {
                c = c1;
                t = 1;
            } else {
                c = c2;
                t = 2;
            }
            cur[i] = -x[(long long)(i - 1) * T + (j - 1)] + c;
            tr[(long long)i * tw + j] = t;
        }
        if (tid == 0) cur[0] = INFINITY;  // cost[0][d] (d >= 1)
        __syncthreads();
    }
    if (tid == 0) {
        int i = N, j = M;
        while (i > 0 || j > 0) {
            if (i >= 1 && j >= 1) jump[(long long)b * n_steps + (i - 1)] = j - 1;  // overwritten until the row's first (smallest) frame
            const int t = (i == 0) ? 2 : (j == 0) ? 1 : tr[(long long)i * tw + j];
            if (t == 0) {
                --i;
                --j;
            } else if (t == 1) {
                --i;
            } else {
                --j;
            }
        }
    }
}

int align_init() {
    cudaError_t e = cudaFuncSetAttribute(align_matrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return set_error("align attr: %s", cudaGetErrorString(e));
    return 0;
}

size_t align_workspace_bytes(int B, int n_sel, int n_steps, int T) {
    return (size_t)B * n_sel * n_steps * 2 * sizeof(float) + 256 + (size_t)B * (kAlignMaxRows + 1) * (T + 1) + 256;
}

int launch_align(const __half* qk, int B, int n_sel, int n_steps, int T, const int* n_tok, const int* row_begin, const int* n_rows,
                 const int* n_frames2, int medfilt, float* matrix, int* jump, void* workspace, size_t ws_bytes, cudaStream_t s) {
    if (n_steps > kAlignMaxRows) return set_error("align: %d token rows > %d", n_steps, kAlignMaxRows);
    if (medfilt != 7 && medfilt != 1) return set_error("align: medfilt_width %d unsupported (7, or 1 = off)", medfilt);
    if (ws_bytes < align_workspace_bytes(B, n_sel, n_steps, T)) return set_error("align: workspace too small");
    float* stat = reinterpret_cast<float*>(workspace);
    unsigned char* trace = reinterpret_cast<unsigned char*>(workspace) + (((size_t)B * n_sel * n_steps * 2 * sizeof(float) + 255) & ~size_t(255));
    {
        const int warps = n_sel * n_steps;
        dim3 grid((warps * 32 + 255) / 256, B);
        align_rowstat_kernel<<<grid, 256, 0, s>>>(qk, stat, n_tok, n_frames2, n_sel, n_steps, T);
        WJB_CHECK_LAUNCH("align_rowstat");
    }
    {
        const size_t smem = ((size_t)n_steps * (kAlignTileW + kAlignStrip) + 2 * kAlignTileW) * sizeof(float);
        if (smem > 200 * 1024) return set_error("align: %d token rows need %zu B of shared memory", n_steps, smem);
        dim3 grid((T + kAlignStrip - 1) / kAlignStrip, B);
        align_matrix_kernel<<<grid, 256, smem, s>>>(qk, stat, matrix, n_tok, n_frames2, n_sel, n_steps, T, medfilt);
        WJB_CHECK_LAUNCH("align_matrix");
    }
    align_dtw_kernel<<<B, kAlignMaxRows, 0, s>>>(matrix, trace, jump, row_begin, n_rows, n_frames2, n_steps, T);
    WJB_CHECK_LAUNCH("align_dtw");
    return 0;
}

}  // namespace wjb
