// Voice-activity gate: Silero-class conv + LSTM stack, one speech probability per 512-sample window.
//
// Replaces the per-window model calls inside get_speech_timestamps(tensor, jit_model) at
// whisperjav/modules/speech_segmentation/backends/silero.py:269-273 / silero_v6.py:205-210 (and the
// per-hop TenVad.process loop at backends/ten.py:232-239), which the reference runs on the CPU one
// window at a time.  The published architecture is restated in whisperjav_b200/vad.py (weights are
// synthetic: the JIT/ONNX files cannot be fetched offline; parity is against oracle/vad_oracle.py).
//
//   u = [64 ctx | 512 samples | 64 reflected]             (640 samples per window)
//   STFT  : 4 steps of a 256-tap Hann DFT (hop 128) -> magnitude [129][4]
//   conv1 : 129->128 k3 s1 ReLU [128][4]   conv2 : 128->64 k3 s2 ReLU [64][2]
//   conv3 : 64->64  k3 s2 ReLU [64][1]     conv4 : 64->128 k3 s1 ReLU [128][1]
//   LSTM  : 128 -> 128 (state carried across windows);  prob = sigmoid(w_out . relu(h) + b)
//
// Two phases.  vad_features_kernel is data-parallel over (clip, 8-window tile): every layer is a small
// GEMM "threads = output channels, registers = (window, step) columns", inputs broadcast from smem,
// weights stored K-major-transposed so a warp reads 128 B per k.  It ends with the input half of the
// LSTM gates (W_ih x + b).  vad_lstm_kernel then walks each clip's windows sequentially: one CTA per
// clip, 512 threads = 512 gate rows, W_hh held in registers, two barriers per window.
#include "../../include/wjb200.h"
#include "kernels.h"

namespace wjb {

constexpr int kWin = 512, kCtx = 64, kU = 640, kTile = 8;  // windows per feature CTA
constexpr int kFeatThreads = 256;

// offsets (floats) into the fp32 weight blob
struct VadOff {
    static constexpr int dft = 0;                         // [256][258]
    static constexpr int c1w = dft + 256 * 258;           // [387][128]
    static constexpr int c1b = c1w + 387 * 128;
    static constexpr int c2w = c1b + 128;                 // [384][64]
    static constexpr int c2b = c2w + 384 * 64;
    static constexpr int c3w = c2b + 64;                  // [192][64]
    static constexpr int c3b = c3w + 192 * 64;
    static constexpr int c4w = c3b + 64;                  // [192][128]
    static constexpr int c4b = c4w + 192 * 128;
    static constexpr int wih = c4b + 128;                 // [128][512]
    static constexpr int bg = wih + 128 * 512;            // [512] (b_ih + b_hh)
    static constexpr int whh = bg + 512;                  // [128][512]
    static constexpr int wout = whh + 128 * 512;          // [128]
    static constexpr int bout = wout + 128;               // [1]
    static constexpr int total = bout + 4;
};

// Y[n][m] = act( b[n] + sum_k X[k][m] * Wt[k][n] ), n over threads, m in registers (M columns).
template <int M, bool kRelu>
__device__ __forceinline__ void smem_linear(const float* __restrict__ X /*smem [K][M]*/, const float* __restrict__ Wt /*global [K][N]*/,
                                            const float* __restrict__ bias, float* __restrict__ Y /*smem [N][M]*/, int K, int N) {
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float acc[M];
        const float b = bias ? __ldg(bias + n) : 0.f;
#pragma unroll
        for (int m = 0; m < M; ++m) acc[m] = b;
        for (int k = 0; k < K; ++k) {
            const float w = __ldg(Wt + (size_t)k * N + n);
            const float4* xr = reinterpret_cast<const float4*>(X + (size_t)k * M);
#pragma unroll
            for (int m4 = 0; m4 < M / 4; ++m4) {
                const float4 x = xr[m4];
                acc[4 * m4 + 0] = fmaf(x.x, w, acc[4 * m4 + 0]);
                acc[4 * m4 + 1] = fmaf(x.y, w, acc[4 * m4 + 1]);
                acc[4 * m4 + 2] = fmaf(x.z, w, acc[4 * m4 + 2]);
                acc[4 * m4 + 3] = fmaf(x.w, w, acc[4 * m4 + 3]);
            }
        }
#pragma unroll
        for (int m = 0; m < M; ++m) Y[(size_t)n * M + m] = kRelu ? fmaxf(acc[m], 0.f) : acc[m];
    }
}

struct VadSmem {
    float buf0[387 * 32];   // largest im2col: conv1 [387][32]  (also the STFT input [256][32])
    float buf1[258 * 32];   // largest output: DFT [258][32]
};

__global__ void __launch_bounds__(kFeatThreads) vad_features_kernel(const float* __restrict__ audio, long long audio_stride,
                                                                    const int* __restrict__ n_samples, const float* __restrict__ W,
                                                                    float* __restrict__ gx /*[clip][n_windows][512]*/, int n_windows) {
    extern __shared__ uint8_t smem_raw[];
    VadSmem& sm = *reinterpret_cast<VadSmem*>(smem_raw);
    const int clip = blockIdx.y, w0 = blockIdx.x * kTile, tid = threadIdx.x;
    const int ns = n_samples[clip];
    const float* a = audio + (long long)clip * audio_stride;
    // ---- STFT input X[k = tap 0..255][m = (w, s)], s = 0..3 : u[128 s + k]
    for (int i = tid; i < 256 * 32; i += kFeatThreads) {
        const int k = i / 32, m = i % 32, w = m / 4, s = m % 4;
        int j = 128 * s + k;                       // index into u (0..639)
        long long g;
        if (j < kCtx + kWin) {
            g = (long long)(w0 + w) * kWin - kCtx + j;
        } else {                                   // reflect the tail of the window: u[576 + r] = x[510 - r]
            const int r = j - (kCtx + kWin);
            g = (long long)(w0 + w) * kWin + (kWin - 2 - r);
        }
        sm.buf0[i] = (g >= 0 && g < ns) ? a[g] : 0.f;
    }
    __syncthreads();
    smem_linear<32, false>(sm.buf0, W + VadOff::dft, nullptr, sm.buf1, 256, 258);  // re: rows 0..128, im: rows 129..257
    __syncthreads();
    // magnitude -> buf0 as conv1 im2col [k = tap*129 + c][m = (w, s)] with zero padding in time
    for (int i = tid; i < 387 * 32; i += kFeatThreads) {
        const int k = i / 32, m = i % 32, w = m / 4, s = m % 4, tap = k / 129, c = k % 129, t = s + tap - 1;
        float v = 0.f;
        if (t >= 0 && t < 4) {
            const float re = sm.buf1[(size_t)c * 32 + w * 4 + t], im = sm.buf1[(size_t)(129 + c) * 32 + w * 4 + t];
            v = sqrtf(re * re + im * im);
        }
        sm.buf0[i] = v;
    }
    __syncthreads();
    smem_linear<32, true>(sm.buf0, W + VadOff::c1w, W + VadOff::c1b, sm.buf1, 387, 128);  // [128][32]
    __syncthreads();
    // conv2 im2col [k = tap*128 + c][m = (w, s2)], s2 = 0..1, stride 2: t = 2 s2 + tap - 1
    for (int i = tid; i < 384 * 16; i += kFeatThreads) {
        const int k = i / 16, m = i % 16, w = m / 2, s = m % 2, tap = k / 128, c = k % 128, t = 2 * s + tap - 1;
        sm.buf0[i] = (t >= 0 && t < 4) ? sm.buf1[(size_t)c * 32 + w * 4 + t] : 0.f;
    }
    __syncthreads();
    smem_linear<16, true>(sm.buf0, W + VadOff::c2w, W + VadOff::c2b, sm.buf1, 384, 64);   // [64][16]
    __syncthreads();
    // conv3 im2col [k = tap*64 + c][m = w], stride 2, one output step: t = tap - 1
    for (int i = tid; i < 192 * 8; i += kFeatThreads) {
        const int k = i / 8, w = i % 8, tap = k / 64, c = k % 64, t = tap - 1;
        sm.buf0[i] = (t >= 0 && t < 2) ? sm.buf1[(size_t)c * 16 + w * 2 + t] : 0.f;
    }
    __syncthreads();
    smem_linear<8, true>(sm.buf0, W + VadOff::c3w, W + VadOff::c3b, sm.buf1, 192, 64);    // [64][8]
    __syncthreads();
    // conv4 im2col: a single time step, so only the centre tap sees data
    for (int i = tid; i < 192 * 8; i += kFeatThreads) {
        const int k = i / 8, w = i % 8, tap = k / 64, c = k % 64;
        sm.buf0[i] = (tap == 1) ? sm.buf1[(size_t)c * 8 + w] : 0.f;
    }
    __syncthreads();
    smem_linear<8, true>(sm.buf0, W + VadOff::c4w, W + VadOff::c4b, sm.buf1, 192, 128);   // x: [128][8]
    __syncthreads();
    smem_linear<8, false>(sm.buf1, W + VadOff::wih, W + VadOff::bg, sm.buf0, 128, 512);   // gx: [512][8]
    __syncthreads();
    for (int i = tid; i < 512 * kTile; i += kFeatThreads) {
        const int w = i / 512, n = i % 512;
        if (w0 + w < n_windows) gx[((long long)clip * n_windows + w0 + w) * 512 + n] = sm.buf0[(size_t)n * 8 + w];
    }
}

// One CTA per clip; thread r owns gate row r (PyTorch order i, f, g, o) of W_hh in registers.
__global__ void __launch_bounds__(512) vad_lstm_kernel(const float* __restrict__ gx, const float* __restrict__ W,
                                                       const int* __restrict__ n_samples, float* __restrict__ probs, int n_windows) {
    const int clip = blockIdx.x, r = threadIdx.x;
    __shared__ float h[128];
    __shared__ float gates[512];
    __shared__ float red[4];
    float wrow[128];
#pragma unroll
    for (int k = 0; k < 128; ++k) wrow[k] = __ldg(W + VadOff::whh + (size_t)k * 512 + r);
    const float wo = r < 128 ? __ldg(W + VadOff::wout + r) : 0.f;
    const float bo = __ldg(W + VadOff::bout);
    float c = 0.f;
    if (r < 128) h[r] = 0.f;
    __syncthreads();
    const int valid = min(n_windows, (n_samples[clip] + kWin - 1) / kWin);
    const float* g = gx + (long long)clip * n_windows * 512;
    for (int t = 0; t < n_windows; ++t) {
        if (t >= valid) {  // past the audio: no speech by definition
            if (r == 0) probs[(long long)clip * n_windows + t] = 0.f;
            continue;
        }
        float acc = g[(long long)t * 512 + r];
#pragma unroll
        for (int k = 0; k < 128; k += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(&h[k]);
            acc = fmaf(wrow[k], hv.x, acc);
            acc = fmaf(wrow[k + 1], hv.y, acc);
            acc = fmaf(wrow[k + 2], hv.z, acc);
            acc = fmaf(wrow[k + 3], hv.w, acc);
        }
        gates[r] = acc;
        __syncthreads();
        float contrib = 0.f;
        if (r < 128) {
            const float ig = 1.f / (1.f + __expf(-gates[r]));
            const float fg = 1.f / (1.f + __expf(-gates[128 + r]));
            const float gg = tanhf(gates[256 + r]);
            const float og = 1.f / (1.f + __expf(-gates[384 + r]));
            c = fg * c + ig * gg;
            const float hn = og * tanhf(c);
            h[r] = hn;
            contrib = wo * fmaxf(hn, 0.f);
            contrib = warp_sum(contrib);
            if ((r & 31) == 0) red[r >> 5] = contrib;
        }
        __syncthreads();
        if (r == 0) probs[(long long)clip * n_windows + t] = 1.f / (1.f + __expf(-(red[0] + red[1] + red[2] + red[3] + bo)));
    }
}

int vad_init() {
    cudaError_t e = cudaFuncSetAttribute(vad_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VadSmem));
    if (e != cudaSuccess) return set_error("vad attr: %s", cudaGetErrorString(e));
    return 0;
}

}  // namespace wjb

using namespace wjb;

extern "C" {

size_t wjb_vad_weights_bytes(void) { return (size_t)VadOff::total * 4; }
size_t wjb_vad_workspace_bytes(int n_clips, int n_windows) { return (size_t)n_clips * n_windows * 512 * 4; }

int wjb_vad_forward(const float* audio, int64_t audio_stride, const int32_t* n_samples, int n_clips, const void* weights, float* probs,
                    int n_windows, void* workspace, void* stream) {
    if (!audio || !n_samples || !weights || !probs || !workspace) return set_error("vad: null argument");
    if (n_clips <= 0 || n_windows <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (int e = ensure_init()) return e;
    float* gx = reinterpret_cast<float*>(workspace);
    dim3 grid((n_windows + kTile - 1) / kTile, n_clips);
    vad_features_kernel<<<grid, kFeatThreads, sizeof(VadSmem), s>>>(audio, audio_stride, n_samples, reinterpret_cast<const float*>(weights), gx,
                                                                   n_windows);
    WJB_CHECK_LAUNCH("vad_features");
    vad_lstm_kernel<<<n_clips, 512, 0, s>>>(gx, reinterpret_cast<const float*>(weights), n_samples, probs, n_windows);
    WJB_CHECK_LAUNCH("vad_lstm");
    return 0;
}

}  // extern "C"
