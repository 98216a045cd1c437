// Micro-benchmark: cost of one tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N, with A read from shared
// memory (SS) or from TMEM (TS).  64 instructions back to back on resident operands, one commit, clock64 around it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I whisperjav_b200/csrc scripts/mma_rate.cu -o scripts/bin/mma_rate
#include "common.cuh"
using namespace wjb;

WJB_DEVINL void umma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a),
                 "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int reps, int per_commit, int wait_each, int mode) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_init(&bar2, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = make_idesc_f16(128, N, 0, 0);
        const uint64_t da = make_smem_desc(smem_u32(smem), 16, 1024, kLayoutSW128);
        const uint64_t db = make_smem_desc(smem_u32(smem + 16384), 16, 1024, kLayoutSW128);
        uint32_t phase = 0;
        // warm-up
        umma_f16(tm, da, db, idesc, 0);
        umma_commit(&bar);
        mbar_wait(&bar, phase);
        phase ^= 1;
        const long long t0 = clock64();
        for (int r = 0; r < reps; r += per_commit) {
            for (int i = 0; i < per_commit; ++i) {
                if (TS)
                    umma_ts(tm, tm + 256 + (i & 3) * 8, db + 2 * (i & 3), idesc, 1);
                else
                    umma_f16(tm, da + 2 * (i & 3), db + 2 * (i & 3), idesc, 1);
            }
            if (mode == 0) umma_commit(&bar);
            if (mode == 1) tc_fence_after();
            if (mode == 2) {
                if (!mbar_try_wait(&bar2, 1)) out[3] = 1;  // an already completed phase: cost of the test itself
            }
            if (mode == 3) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(out[4]));
            if (mode == 0 && per_commit < reps && wait_each) {  // wait per group: exposes the pipeline latency
                mbar_wait(&bar, phase);
                phase ^= 1;
            }
        }
        const long long t1 = clock64();
        if (per_commit >= reps) {
            mbar_wait(&bar, phase);
            phase ^= 1;
        } else if (!wait_each || mode != 0) {
            umma_commit(&bar2);
            mbar_wait(&bar2, 0);
        }
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

template <int N, bool TS>
void run(const char* name, long long* d_out, int reps, int per_commit, int wait_each = 1, int mode = 0) {
    cudaFuncSetAttribute(rate_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    rate_kernel<N, TS><<<1, 128, 60 * 1024>>>(d_out, reps, per_commit, wait_each, mode);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-28s reps %3d per_commit %3d: issue %6lld clk, total %6lld clk -> %.1f clk / MMA  (%s)\n", name, reps, per_commit, h[0], h[1],
           (double)h[1] / reps, cudaGetErrorString(e));
}


// The k loop as the GEMM kernels run it: per group of 4 MMAs fresh descriptors (another pipeline stage).  kUniform = the whole
// warp runs the loop and one elected lane issues (descriptor arithmetic can stay in uniform registers); otherwise a single
// thread inside a divergent branch does everything, as `if (lane == 0)` code does.
template <int N, bool kUniform>
__global__ void __launch_bounds__(128, 1) loop_kernel(long long* out, int groups) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    constexpr uint32_t idesc = make_idesc_f16(128, N, 0, 0);
    if (threadIdx.x < 32) {
        if (kUniform) {
            const long long t0 = clock64();
            int stage = 0;
            for (int g = 0; g < groups; ++g) {
                const uint64_t da = make_smem_desc(smem_u32(smem + stage * 16384), 16, 1024, kLayoutSW128);
                const uint64_t db = make_smem_desc(smem_u32(smem + 65536 + stage * 32768), 16, 1024, kLayoutSW128);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(tm, da + 2 * k, db + 2 * k, idesc, 1);
                }
                __syncwarp();
                if (++stage == 4) stage = 0;
            }
            const long long t1 = clock64();
            if (elect_one()) umma_commit(&bar);
            __syncwarp();
            mbar_wait(&bar, 0);
            const long long t2 = clock64();
            if (threadIdx.x == 0) {
                out[0] = t1 - t0;
                out[1] = t2 - t0;
            }
        } else if (threadIdx.x == 0) {
            const long long t0 = clock64();
            int stage = 0;
            for (int g = 0; g < groups; ++g) {
                const uint64_t da = make_smem_desc(smem_u32(smem + stage * 16384), 16, 1024, kLayoutSW128);
                const uint64_t db = make_smem_desc(smem_u32(smem + 65536 + stage * 32768), 16, 1024, kLayoutSW128);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(tm, da + 2 * k, db + 2 * k, idesc, 1);
                if (++stage == 4) stage = 0;
            }
            const long long t1 = clock64();
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            const long long t2 = clock64();
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

template <int N, bool kUniform>
void run_loop(const char* name, long long* d_out, int groups) {
    cudaFuncSetAttribute(loop_kernel<N, kUniform>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    loop_kernel<N, kUniform><<<1, 128, 200 * 1024>>>(d_out, groups);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-40s groups %3d: issue %6lld clk, total %6lld clk -> %.1f clk / MMA  (%s)\n", name, groups, h[0], h[1], (double)h[1] / (4 * groups),
           cudaGetErrorString(e));
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 64);
    run<64, false>("SS N=64", d_out, 64, 64);
    run<128, false>("SS N=128", d_out, 64, 64);
    run<256, false>("SS N=256", d_out, 64, 64);
    run<64, true>("TS N=64", d_out, 64, 64);
    run<128, true>("TS N=128", d_out, 64, 64);
    run<256, true>("TS N=256", d_out, 64, 64);
    run<64, false>("SS N=64 commit+wait per 4", d_out, 64, 4);
    run<64, true>("TS N=64 commit+wait per 4", d_out, 64, 4);
    run<256, false>("SS N=256 commit+wait per 4", d_out, 64, 4);
    run<64, false>("SS N=64 commit+wait per 1", d_out, 64, 1);
    run<64, false>("SS N=64 commit per 4 no wait", d_out, 64, 4, 0);
    run<256, false>("SS N=256 commit per 4 no wait", d_out, 64, 4, 0);
    run<256, false>("SS N=256 commit per 8 no wait", d_out, 64, 8, 0);
    run<64, false>("SS N=64 fence::after per 4", d_out, 64, 4, 0, 1);
    run<64, false>("SS N=64 try_wait per 4", d_out, 64, 4, 0, 2);
    run<64, false>("SS N=64 globaltimer per 4", d_out, 64, 4, 0, 3);
    run<64, false>("SS N=64 nothing per 4", d_out, 64, 4, 0, 4);
    run_loop<64, false>("k loop N=64, one thread in a branch", d_out, 16);
    run_loop<64, true>("k loop N=64, warp-uniform + elect", d_out, 16);
    run_loop<256, false>("k loop N=256, one thread in a branch", d_out, 16);
    run_loop<256, true>("k loop N=256, warp-uniform + elect", d_out, 16);
    return 0;
}
