// Energy gate of the two-pass scene detector: the per-analysis-window signal energy that `auditok.split()` computes on
// the CPU for every 50 ms block of a whole film (reference: whisperjav/modules/scene_detection_backends/
// auditok_backend.py:379-392 pass 1 over the stream, :552-567 pass 2 over every oversized chapter; upstream auditok
// 0.3.0 signal.calculate_energy / AudioEnergyValidator.is_valid on the int16 bytes the reference builds with
// `(audio * 32767).astype(np.int16)`).
//
// What the device returns is the EXACT integer sum of squares of those int16 samples per window, so the host takes the
// decision `20 log10(sqrt(sum / n)) >= threshold` in float64 exactly as upstream does (the float64 mean of integers below
// 2^53 is exact): the valid/silent flags are bit-identical to the CPU path by construction, not to a tolerance.
//
// HBM-bound integer work: the stream is read once (4 B per sample, 3.2 KB per window), 8 B written per window.  One warp
// per window; a window whose first sample is 16-byte aligned (always the case in pass 1) is read as float4, otherwise
// sample by sample (still one 128-byte line per warp load).  Windows of all regions of a pass share one launch: the
// region of a window is found by bisection over the window prefix counts (a few hundred regions at most).
#include "kernels.h"

namespace wjb {

constexpr int kSceneWarps = 8;

WJB_DEVINL unsigned long long sq_i16(float x) {
    // numpy: float32 array * python int -> float32 multiply; astype(int16) truncates toward zero (through int32, low 16 bits kept)
    const int q = __float2int_rz(x * 32767.0f);
    const long long s = (long long)(short)q;
    return (unsigned long long)(s * s);
}

__global__ void __launch_bounds__(kSceneWarps * 32)
scene_energy_kernel(const float* __restrict__ audio, long long n_audio, const long long* __restrict__ region_start,
                    const long long* __restrict__ region_len, const long long* __restrict__ window_base, int n_regions, int window,
                    unsigned long long* __restrict__ sumsq, long long n_windows) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * kSceneWarps + warp;
    if (w >= n_windows) return;
    int lo = 0, hi = n_regions - 1;  // last region with window_base[r] <= w
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (window_base[mid] <= w) lo = mid; else hi = mid - 1;
    }
    const long long j = w - window_base[lo];
    const long long r0 = region_start[lo], rl = region_len[lo];
    long long begin = r0 + j * window;
    long long end = begin + window;
    if (end > r0 + rl) end = r0 + rl;  // the last block of a region may be short (auditok reads what is left)
    if (end > n_audio) end = n_audio;
    if (begin > end) begin = end;
    unsigned long long acc = 0;
    const float* p = audio + begin;
    const long long n = end - begin;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const long long n4 = n >> 2;
        const float4* p4 = reinterpret_cast<const float4*>(p);
        for (long long i = lane; i < n4; i += 32) {
            const float4 v = __ldg(p4 + i);
            acc += sq_i16(v.x) + sq_i16(v.y) + sq_i16(v.z) + sq_i16(v.w);
        }
        for (long long i = (n4 << 2) + lane; i < n; i += 32) acc += sq_i16(__ldg(p + i));
    } else {
        for (long long i = lane; i < n; i += 32) acc += sq_i16(__ldg(p + i));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) sumsq[w] = acc;
}

}  // namespace wjb

extern "C" {

int wjb_scene_energy(const float* audio, int64_t n_audio, const int64_t* region_start, const int64_t* region_len, const int64_t* window_base,
                     int n_regions, int window, uint64_t* sumsq, int64_t n_windows, void* stream) {
    using namespace wjb;
    if (!audio || !region_start || !region_len || !window_base || !sumsq) return set_error("wjb_scene_energy: null pointer");
    if (n_regions < 1 || window < 1 || n_windows < 0 || n_audio < 0) return set_error("wjb_scene_energy: bad sizes (regions=%d window=%d)", n_regions, window);
    if (n_windows == 0) return 0;
    const long long blocks = (n_windows + kSceneWarps - 1) / kSceneWarps;
    if (blocks > 0x7fffffffLL) return set_error("wjb_scene_energy: too many windows");
    scene_energy_kernel<<<(unsigned)blocks, kSceneWarps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        audio, (long long)n_audio, reinterpret_cast<const long long*>(region_start), reinterpret_cast<const long long*>(region_len),
        reinterpret_cast<const long long*>(window_base), n_regions, window, reinterpret_cast<unsigned long long*>(sumsq), (long long)n_windows);
    WJB_CHECK_LAUNCH("scene_energy_kernel");
    return 0;
}

}  // extern "C"
