// LayerNorm (fp32 statistics, fp16 in/out -- openai-whisper model.py::LayerNorm) and the
// explicit im2col fallback for the k=3 convolutions.
#include "kernels.h"

namespace wjb {

// One warp per row; the row (<= 2048 halfs) lives in registers between the two passes.
template <int kVecPerLane>
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma,
                                                        const __half* __restrict__ beta, __half* __restrict__ out, int rows, int n) {
    pdl_prologue();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int nvec = n >> 3;
    const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)warp * n);
    float v[kVecPerLane][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            uint4 u = xr[idx];
            const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f = __half22float2(h2[j]);
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
    for (int i = 0; i < kVecPerLane; ++i) {
        if (lane + i * 32 < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = v[i][j] - mean;
                sq += d * d;
            }
        }
    }
    sq = warp_sum(sq);
    const float rstd = rsqrtf(sq / n + 1e-5f);
    uint4* orow = reinterpret_cast<uint4*>(out + (size_t)warp * n);
    const uint4* g4 = reinterpret_cast<const uint4*>(gamma);
    const uint4* b4 = reinterpret_cast<const uint4*>(beta);
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            uint4 g = __ldg(g4 + idx), b = __ldg(b4 + idx), o;
            const __half2* gh = reinterpret_cast<const __half2*>(&g);
            const __half2* bh = reinterpret_cast<const __half2*>(&b);
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 gf = __half22float2(gh[j]), bf = __half22float2(bh[j]);
                oh[j] = __floats2half2_rn((v[i][2 * j] - mean) * rstd * gf.x + bf.x, (v[i][2 * j + 1] - mean) * rstd * gf.y + bf.y);
            }
            orow[idx] = o;
        }
    }
}

int launch_layernorm(const __half* x, const __half* gamma, const __half* beta, __half* out, int rows, int n, cudaStream_t s) {
    if (n % 8 != 0 || n > 2048) return set_error("layernorm: n=%d unsupported (multiple of 8, <= 2048)", n);
    if (rows <= 0) return 0;
    const int warps_per_block = 8;
    const int grid = (rows + warps_per_block - 1) / warps_per_block;
    const int nvec = n / 8;
    if (nvec <= 64)
        launch_k(layernorm_kernel<2>, dim3(grid), dim3(256), 0, s, x, gamma, beta, out, rows, n);
    else if (nvec <= 160)
        launch_k(layernorm_kernel<5>, dim3(grid), dim3(256), 0, s, x, gamma, beta, out, rows, n);
    else
        launch_k(layernorm_kernel<8>, dim3(grid), dim3(256), 0, s, x, gamma, beta, out, rows, n);
    WJB_CHECK_LAUNCH("layernorm");
    return 0;
}

// Frame classifier head of the WhisperSeg-class VAD (speech_segmentation/backends/whisperseg.py:355-393: encoder states ->
// one logit per 20 ms frame -> sigmoid): prob[r] = sigmoid(x[r] . w + b), one warp per frame row, fp32 accumulate.
__global__ void __launch_bounds__(256) frame_head_kernel(const __half* __restrict__ x, const __half* __restrict__ w, float bias,
                                                         float* __restrict__ prob, int rows, int n) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)warp * n);
    const uint4* wr = reinterpret_cast<const uint4*>(w);
    float acc = 0.f;
    for (int i = lane; i < (n >> 3); i += 32) {
        const uint4 a = xr[i], b = __ldg(wr + i);
        const __half2* ah = reinterpret_cast<const __half2*>(&a);
        const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 af = __half22float2(ah[j]), bf = __half22float2(bh[j]);
            acc = fmaf(af.x, bf.x, acc);
            acc = fmaf(af.y, bf.y, acc);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) prob[warp] = 1.0f / (1.0f + expf(-(acc + bias)));
}

int launch_frame_head(const __half* x, const __half* w, float bias, float* prob, int rows, int n, cudaStream_t s) {
    if (n % 8) return set_error("frame_head: n must be a multiple of 8");
    if (rows <= 0) return 0;
    frame_head_kernel<<<(rows + 7) / 8, 256, 0, s>>>(x, w, bias, prob, rows, n);
    WJB_CHECK_LAUNCH("frame_head");
    return 0;
}

// out[b][t][k*C + c] = xpad[b][stride*t + k][c]  (k = 0..2), 16-byte vectors along c.
__global__ void im2col_k3_kernel(const uint4* __restrict__ xpad, uint4* __restrict__ out, int B, int T_out, int Cv, int stride,
                                 int T_in_padded) {
    const long long total = (long long)B * T_out * 3 * Cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = i % Cv;
        const int k = (i / Cv) % 3;
        const int t = (i / (3 * Cv)) % T_out;
        const int b = i / ((long long)3 * Cv * T_out);
        out[i] = xpad[((long long)b * T_in_padded + stride * t + k) * Cv + c];
    }
}

int launch_im2col_k3(const __half* xpad, __half* out, int B, int T_out, int C, int stride, int T_in_padded, cudaStream_t s) {
    if (C % 8) return set_error("im2col: C must be a multiple of 8");
    im2col_k3_kernel<<<sm_count() * 8, 256, 0, s>>>(reinterpret_cast<const uint4*>(xpad), reinterpret_cast<uint4*>(out), B, T_out, C / 8,
                                                  stride, T_in_padded);
    WJB_CHECK_LAUNCH("im2col_k3");
    return 0;
}

}  // namespace wjb
