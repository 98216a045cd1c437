// Fused STFT(400, hop 160, periodic Hann, centred/reflect) + |.|^2 + Slaney mel + log10 + scale.
//
// Replaces openai-whisper audio.py::log_mel_spectrogram (torch.stft -> cuFFT, filters @ mag ->
// cuBLAS, five element-wise kernels), reached from whisperjav/modules/whisper_pro_asr.py:433 and,
// for the HF path, WhisperFeatureExtractor at generators/anime_whisper.py:256-263.
//
// One warp transforms one frame with an in-register/in-smem mixed-radix FFT, N = 400 = 16 x 25:
//   n = 25*n1 + n2, k = k1 + 16*k2
//   step 1 (lane = n2 < 25): 16-point real DFT over n1 -> Y[k1], k1 = 0..8, times W400^(n2*k1)
//   step 2 (lane = k2 < 25): X[k1 + 16*k2] = sum_n2 Y[k1][n2] * W25^(n2*k2)   (W25 row in registers)
// Only 201 bins are needed; bins with k > 200 fold back by conjugate symmetry (|X[400-k]| = |X[k]|).
// Output is half((log10(max(mel,1e-10)) + 4) / 4) plus the per-clip fp32 max of log10; the
// "max - 8" floor is applied by logmel_floor_kernel (max commutes with the monotone fp16 rounding,
// so there is no double rounding).  Frames >= n_samples/160 are literal zeros (pad_or_trim).
#include "kernels.h"

namespace wjb {

constexpr int kNFFT = 400;
constexpr int kHop = 160;
constexpr int kNFreq = 201;
constexpr int kFT = 32;  // frames per CTA
constexpr int kMelThreads = 256;
constexpr int kSpan = kFT * kHop + (kNFFT - kHop);  // 5360 samples staged per CTA

__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

struct MelSmem {
    float samples[kSpan];
    float window[kNFFT];
    float2 w400[200];
    float2 Y[8][9][25];
    float P[8][208];
    int mel_lo[128], mel_hi[128];
    __half tile[kFT][128];
};

__global__ void __launch_bounds__(kMelThreads) logmel_kernel(const LogmelArgs a) {
    extern __shared__ uint8_t smem_raw[];
    MelSmem& sm = *reinterpret_cast<MelSmem*>(smem_raw);
    const int clip = blockIdx.y;
    const int t0 = blockIdx.x * kFT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_samples = a.n_samples[clip];
    const int content_frames = n_samples / kHop;
    const int max_frames = (n_samples + kNFFT / 2 + kHop - 1) / kHop;  // frames whose window still touches audio
    const float* audio = a.audio + (long long)clip * a.audio_stride;

    if (t0 >= max_frames && t0 >= content_frames) {
        // whole tile is padding: literal zeros
        if (a.time_major) {
            for (int i = tid; i < kFT * a.n_mels; i += kMelThreads) {
                const int f = i / a.n_mels, m = i % a.n_mels;
                if (t0 + f < a.n_frames) a.out[(long long)clip * a.out_clip_stride + (long long)(a.row0 + t0 + f) * a.n_mels + m] = __float2half(0.f);
            }
        } else {
            for (int i = tid; i < kFT * a.n_mels; i += kMelThreads) {
                const int m = i / kFT, f = i % kFT;
                if (t0 + f < a.n_frames) a.out[(long long)clip * a.out_clip_stride + (long long)m * a.n_frames + t0 + f] = __float2half(0.f);
            }
        }
        return;
    }

    // ---- stage samples (reflect at 0, optional reflect at the padded end, zeros past the audio)
    for (int i = tid; i < kSpan; i += kMelThreads) {
        long long g = (long long)t0 * kHop - kNFFT / 2 + i;
        if (g < 0) g = -g;
        if (a.reflect_total > 0 && g >= a.reflect_total) g = 2LL * (a.reflect_total - 1) - g;
        sm.samples[i] = (g >= 0 && g < n_samples) ? audio[g] : 0.f;
    }
    for (int i = tid; i < kNFFT; i += kMelThreads) sm.window[i] = (float)(0.5 - 0.5 * cospi(2.0 * i / kNFFT));
    for (int i = tid; i < 200; i += kMelThreads) {
        double s, c;
        sincospi(-2.0 * i / 400.0, &s, &c);
        sm.w400[i] = make_float2((float)c, (float)s);
    }
    for (int m = tid; m < a.n_mels; m += kMelThreads) {
        sm.mel_lo[m] = a.mel_range[2 * m];
        sm.mel_hi[m] = a.mel_range[2 * m + 1];
    }
    // W25 row of this lane (k2 = lane): w[n2] = exp(-2 pi i n2 k2 / 25)
    float wr[25], wi[25];
    {
        const int k2 = lane < 25 ? lane : 0;
#pragma unroll
        for (int n2 = 0; n2 < 25; ++n2) {
            float s, c;
            sincospif(-2.0f * (float)((n2 * k2) % 25) / 25.0f, &s, &c);
            wr[n2] = c;
            wi[n2] = s;
        }
    }
    __syncthreads();

    // cos/sin(2 pi m / 16)
    const float C16[16] = {1.f, 0.92387953251f, 0.70710678119f, 0.38268343237f, 0.f, -0.38268343237f, -0.70710678119f, -0.92387953251f,
                           -1.f, -0.92387953251f, -0.70710678119f, -0.38268343237f, 0.f, 0.38268343237f, 0.70710678119f, 0.92387953251f};
    const float S16[16] = {0.f, 0.38268343237f, 0.70710678119f, 0.92387953251f, 1.f, 0.92387953251f, 0.70710678119f, 0.38268343237f,
                           0.f, -0.38268343237f, -0.70710678119f, -0.92387953251f, -1.f, -0.92387953251f, -0.70710678119f, -0.38268343237f};

    float local_max = -INFINITY;
    for (int fi = 0; fi < kFT / 8; ++fi) {
        const int f = warp * (kFT / 8) + fi;
        const int t = t0 + f;
        const bool compute = t < max_frames;       // frame touches audio -> contributes to the clip max
        const bool writes = t < content_frames;    // frame is part of the returned content
        if (compute) {
            // ---- step 1
            if (lane < 25) {
                const int n2 = lane;
                float e[8], o[8];
#pragma unroll
                for (int n1 = 0; n1 < 8; ++n1) {
                    const int i0 = 25 * n1 + n2, i1 = 25 * (n1 + 8) + n2;
                    const float x0 = sm.samples[f * kHop + i0] * sm.window[i0];
                    const float x1 = sm.samples[f * kHop + i1] * sm.window[i1];
                    e[n1] = x0 + x1;
                    o[n1] = x0 - x1;
                }
#pragma unroll
                for (int k1 = 0; k1 <= 8; ++k1) {
                    float yr = 0.f, yi = 0.f;
#pragma unroll
                    for (int n1 = 0; n1 < 8; ++n1) {
                        const float v = (k1 & 1) ? o[n1] : e[n1];
                        const int m = (n1 * k1) & 15;
                        yr = fmaf(v, C16[m], yr);
                        yi = fmaf(v, -S16[m], yi);
                    }
                    const float2 w = sm.w400[n2 * k1];
                    sm.Y[warp][k1][n2] = make_float2(yr * w.x - yi * w.y, yr * w.y + yi * w.x);
                }
            }
            __syncwarp();
            // ---- step 2
            if (lane < 25) {
                const int k2 = lane;
#pragma unroll 1
                for (int k1 = 0; k1 <= 8; ++k1) {
                    if ((k1 == 0 || k1 == 8) && k2 > 12) continue;
                    float xr = 0.f, xi = 0.f;
#pragma unroll
                    for (int n2 = 0; n2 < 25; ++n2) {
                        const float2 y = sm.Y[warp][k1][n2];
                        xr = fmaf(y.x, wr[n2], xr);
                        xr = fmaf(-y.y, wi[n2], xr);
                        xi = fmaf(y.x, wi[n2], xi);
                        xi = fmaf(y.y, wr[n2], xi);
                    }
                    int kk = k1 + 16 * k2;
                    if (kk > 200) kk = 400 - kk;
                    sm.P[warp][kk] = xr * xr + xi * xi;
                }
            }
            __syncwarp();
            // ---- mel + log
            for (int m = lane; m < a.n_mels; m += 32) {
                const int lo = sm.mel_lo[m], hi = sm.mel_hi[m];
                float acc = 0.f;
                for (int k = lo; k < hi; ++k) acc = fmaf(__ldg(a.filters + m * kNFreq + k), sm.P[warp][k], acc);
                const float lg = log10f(fmaxf(acc, 1e-10f));
                local_max = fmaxf(local_max, lg);
                sm.tile[f][m] = __float2half_rn(writes ? (lg + 4.0f) / 4.0f : 0.f);
            }
            __syncwarp();
        } else {
            for (int m = lane; m < a.n_mels; m += 32) sm.tile[f][m] = __float2half(0.f);
        }
    }
    local_max = warp_max(local_max);
    if (lane == 0 && local_max > -INFINITY) atomicMax(reinterpret_cast<unsigned*>(a.clip_max) + clip, f2ord(local_max));
    __syncthreads();

    // ---- store the [kFT][n_mels] tile
    if (a.time_major) {
        const int vec_per_row = a.n_mels / 8;
        for (int i = tid; i < kFT * vec_per_row; i += kMelThreads) {
            const int f = i / vec_per_row, v = i % vec_per_row;
            if (t0 + f < a.n_frames) {
                uint4* dst = reinterpret_cast<uint4*>(a.out + (long long)clip * a.out_clip_stride + (long long)(a.row0 + t0 + f) * a.n_mels) + v;
                *dst = *reinterpret_cast<const uint4*>(&sm.tile[f][v * 8]);
            }
        }
    } else {
        for (int i = tid; i < kFT * a.n_mels; i += kMelThreads) {
            const int m = i / kFT, f = i % kFT;
            if (t0 + f < a.n_frames) a.out[(long long)clip * a.out_clip_stride + (long long)m * a.n_frames + t0 + f] = sm.tile[f][m];
        }
    }
}

// Non-zero span [lo, hi) of every mel filter row (run once per launch; 1 CTA).
__global__ void mel_range_kernel(const float* __restrict__ filters, int n_mels, int* __restrict__ range) {
    for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
        int lo = kNFreq, hi = 0;
        for (int k = 0; k < kNFreq; ++k) {
            if (filters[m * kNFreq + k] != 0.f) {
                lo = min(lo, k);
                hi = max(hi, k + 1);
            }
        }
        range[2 * m] = lo;
        range[2 * m + 1] = hi;
    }
}

// y = max(y, half(((max - 8) + 4) / 4)) over the content frames of each clip.
__global__ void logmel_floor_kernel(const LogmelArgs a) {
    const int clip = blockIdx.y;
    const int content_frames = min(a.n_samples[clip] / kHop, a.n_frames);
    const unsigned ord = reinterpret_cast<const unsigned*>(a.clip_max)[clip];
    if (ord == 0u) return;  // no frame computed
    const float gmax = ord2f(ord);
    const __half flo = __float2half_rn(((gmax - 8.0f) + 4.0f) / 4.0f);
    const long long total = (long long)content_frames * a.n_mels;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long off;
        if (a.time_major) {
            off = (long long)clip * a.out_clip_stride + (long long)a.row0 * a.n_mels + i;
        } else {
            const int m = i / content_frames, t = i % content_frames;
            off = (long long)clip * a.out_clip_stride + (long long)m * a.n_frames + t;
        }
        const __half v = a.out[off];
        if (__hlt(v, flo)) a.out[off] = flo;
    }
}

int launch_logmel(const LogmelArgs& a, cudaStream_t s) {
    if (a.n_mels > 128 || a.n_mels % 8) return set_error("logmel: n_mels=%d unsupported (<=128, multiple of 8)", a.n_mels);
    if (a.n_clips <= 0 || a.n_frames <= 0) return 0;
    cudaError_t e = cudaMemsetAsync(a.clip_max, 0, sizeof(float) * a.n_clips, s);
    if (e != cudaSuccess) return set_error("logmel memset: %s", cudaGetErrorString(e));
    mel_range_kernel<<<1, 128, 0, s>>>(a.filters, a.n_mels, a.mel_range);
    WJB_CHECK_LAUNCH("mel_range");
    dim3 grid((a.n_frames + kFT - 1) / kFT, a.n_clips);
    logmel_kernel<<<grid, kMelThreads, sizeof(MelSmem), s>>>(a);
    WJB_CHECK_LAUNCH("logmel");
    dim3 grid2(64, a.n_clips);
    logmel_floor_kernel<<<grid2, 256, 0, s>>>(a);
    WJB_CHECK_LAUNCH("logmel_floor");
    return 0;
}

int logmel_init() {
    cudaError_t e = cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MelSmem));
    if (e != cudaSuccess) return set_error("logmel attr: %s", cudaGetErrorString(e));
    return 0;
}

}  // namespace wjb
