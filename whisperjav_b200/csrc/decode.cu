// Decoder token logic on device: embedding lookup and the fused logit filters + greedy update.
//
// sample_kernel restates openai-whisper decoding.py for one step, per batch row, entirely on device:
//   SuppressBlank, SuppressTokens, ApplyTimestampRules (pairs, monotonicity, initial-timestamp
//   constraint, "timestamp mass > max text prob"), GreedyDecoder.update (argmax at T == 0,
//   sum_logprobs += logprob while the row is alive, EOT latch), and at the SOT position the
//   no_speech probability  softmax(logits)[no_speech].
// The host reads tokens / sum_logprob / no_speech_prob back once per decode run, not per token.
#include "kernels.h"

namespace wjb {

__global__ void embed_kernel(const int* __restrict__ tokens, int tokens_stride, const __half* __restrict__ emb,
                             const __half* __restrict__ pos, __half* __restrict__ x, const DecodeCtl* __restrict__ ctl, int n) {
    pdl_prologue();
    const int b = blockIdx.x;
    const int step = ctl->step;
    const int tok = tokens[(long long)b * tokens_stride + step];
    const __half2* e = reinterpret_cast<const __half2*>(emb + (long long)tok * n);
    const __half2* p = reinterpret_cast<const __half2*>(pos + (long long)step * n);
    __half2* o = reinterpret_cast<__half2*>(x + (long long)b * n);
    for (int i = threadIdx.x; i < n / 2; i += blockDim.x) {
        const float2 a = __half22float2(e[i]), c = __half22float2(p[i]);
        o[i] = __floats2half2_rn(a.x + c.x, a.y + c.y);
    }
}

int launch_embed(const int* tokens, int tokens_stride, const __half* emb, const __half* pos, __half* x, const DecodeCtl* ctl, int B,
                 int n, cudaStream_t s) {
    launch_k(embed_kernel, dim3(B), dim3(128), 0, s, tokens, tokens_stride, emb, pos, x, ctl, n);
    WJB_CHECK_LAUNCH("embed");
    return 0;
}

constexpr int kSampleThreads = 1024;

struct ArgMax {
    float v;
    int i;
};
__device__ __forceinline__ ArgMax amax(ArgMax a, ArgMax b) {
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}
__device__ __forceinline__ ArgMax block_argmax(ArgMax a, ArgMax* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgMax b;
        b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
        b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
        a = amax(a, b);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = a;
    __syncthreads();
    ArgMax r = sh[0];
    for (int w = 1; w < kSampleThreads / 32; ++w) r = amax(r, sh[w]);
    return r;
}
__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = 0.f;
    for (int w = 0; w < kSampleThreads / 32; ++w) r += sh[w];
    return r;
}

__global__ void __launch_bounds__(kSampleThreads)
sample_kernel(const __half* __restrict__ logits, const unsigned char* __restrict__ suppress_mask, int* __restrict__ tokens,
              float* __restrict__ sum_logprob, float* __restrict__ no_speech_prob, int* __restrict__ out_len,
              unsigned char* __restrict__ done, DecodeCtl* __restrict__ ctl, const DecodeParams p) {
    extern __shared__ __half srow[];
    __shared__ ArgMax sh_am[kSampleThreads / 32];
    __shared__ float sh_f[kSampleThreads / 32];
    __shared__ int st[4];  // last_was_ts, penult_was_ts, have_ts, timestamp_last
    pdl_prologue();
    const int b = blockIdx.x, tid = threadIdx.x;
    const int step = ctl->step;
    const int cur_len = step + 1;
    const int n_initial = ctl->n_initial;
    const bool at_sot = (step == ctl->sot_index);
    const bool sampling = cur_len >= n_initial;
    if (!at_sot && (!sampling || done[b])) return;
    const int V = p.n_vocab;
    const __half* row = logits + (long long)b * p.logits_stride;
    for (int v = tid; v < V; v += kSampleThreads) srow[v] = row[v];
    int* trow = tokens + (long long)b * p.tokens_stride;
    if (tid == 0) {
        int last_ts = 0, pen_ts = 0, have = 0, tl = 0;
        const int ns = cur_len - n_initial;  // sampled so far
        if (sampling && p.apply_timestamp_rules) {
            last_ts = ns >= 1 && trow[cur_len - 1] >= p.timestamp_begin;
            pen_ts = ns < 2 || trow[cur_len - 2] >= p.timestamp_begin;
            for (int i = cur_len - 1; i >= n_initial; --i) {
                if (trow[i] >= p.timestamp_begin) {
                    have = 1;
                    tl = trow[i];
                    break;
                }
            }
            if (have && !(last_ts && !pen_ts)) tl += 1;
        }
        st[0] = last_ts;
        st[1] = pen_ts;
        st[2] = have;
        st[3] = tl;
    }
    __syncthreads();

    if (at_sot) {
        ArgMax a{-INFINITY, 0};
        for (int v = tid; v < V; v += kSampleThreads) a = amax(a, ArgMax{__half2float(srow[v]), v});
        a = block_argmax(a, sh_am);
        float s = 0.f;
        for (int v = tid; v < V; v += kSampleThreads) s += __expf(__half2float(srow[v]) - a.v);
        s = block_sum(s, sh_f);
        if (tid == 0) no_speech_prob[b] = __expf(__half2float(srow[p.no_speech]) - a.v) / s;
    }
    if (!sampling || done[b]) return;

    const bool first = (cur_len == n_initial);
    const int last_ts = st[0], pen_ts = st[1], have_ts = st[2], ts_last = st[3];
    const int tsb = p.timestamp_begin;
    auto masked = [&](int v) -> bool {
        if (p.suppress_blank && first && (v == p.blank_token || v == p.eot)) return true;
        if (suppress_mask && suppress_mask[v]) return true;
        if (p.apply_timestamp_rules) {
            if (v == p.no_timestamps) return true;
            if (last_ts) {
                if (pen_ts) {
                    if (v >= tsb) return true;
                } else {
                    if (v < p.eot) return true;
                }
            }
            if (have_ts && v >= tsb && v < ts_last) return true;
            if (first) {
                if (v < tsb) return true;
                if (p.max_initial_timestamp_index >= 0 && v > tsb + p.max_initial_timestamp_index) return true;
            }
        }
        return false;
    };
    // pass 1: masked maxima of the text part [0, tsb) and the timestamp part [tsb, V)
    const int split = p.apply_timestamp_rules ? tsb : V;
    ArgMax at{-INFINITY, 0x7fffffff}, as{-INFINITY, 0x7fffffff};
    for (int v = tid; v < V; v += kSampleThreads) {
        if (masked(v)) {
            srow[v] = __float2half(-INFINITY);
            continue;
        }
        const float l = __half2float(srow[v]);
        if (v < split)
            at = amax(at, ArgMax{l, v});
        else
            as = amax(as, ArgMax{l, v});
    }
    at = block_argmax(at, sh_am);
    as = block_argmax(as, sh_am);
    const float mx = fmaxf(at.v, as.v);
    float s_text = 0.f, s_ts = 0.f;
    for (int v = tid; v < V; v += kSampleThreads) {
        const float l = __half2float(srow[v]);
        if (l == -INFINITY) continue;
        const float e = expf(l - mx);
        if (v < split)
            s_text += e;
        else
            s_ts += e;
    }
    s_text = block_sum(s_text, sh_f);
    s_ts = block_sum(s_ts, sh_f);
    bool ts_wins = false;
    if (p.apply_timestamp_rules && s_ts > 0.f) {
        // logsumexp(timestamp logprobs) > max text logprob  <=>  log(s_ts) + mx > max text logit
        ts_wins = (logf(s_ts) + mx) > at.v;
    }
    ArgMax pick = ts_wins ? as : amax(at, as);
    const float temperature = ctl->temperature;
    if (temperature > 0.f) {
        // GreedyDecoder.update at T > 0: Categorical(logits / T) over the filtered logits, drawn by Gumbel-max with a
        // counter-based generator keyed on (seed, row, step, token id); logprobs below stay those of the unscaled logits
        const float inv_t = 1.0f / temperature;
        const unsigned seed = ctl->seed;
        ArgMax g{-INFINITY, 0x7fffffff};
        for (int v = tid; v < V; v += kSampleThreads) {
            const float l = __half2float(srow[v]);
            if (l == -INFINITY || (ts_wins && v < split)) continue;
            unsigned hsh = seed ^ (0x9E3779B9u * (unsigned)(b + 1)) ^ (0x85EBCA6Bu * (unsigned)(step + 1)) ^ (0xC2B2AE35u * (unsigned)(v + 1));
            hsh ^= hsh >> 16;
            hsh *= 0x7feb352du;
            hsh ^= hsh >> 15;
            hsh *= 0x846ca68bu;
            hsh ^= hsh >> 16;
            const float u = (float)(hsh >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f);  // (0, 1)
            g = amax(g, ArgMax{l * inv_t - logf(-logf(u)), v});
        }
        pick = block_argmax(g, sh_am);
    }
    if (tid == 0) {
        const int tok = pick.i;
        const float lse = ts_wins ? mx + logf(s_ts) : mx + logf(s_text + s_ts);
        const float lp = __half2float(srow[tok]) - lse;
        sum_logprob[b] += lp;
        if (cur_len < p.tokens_stride) trow[cur_len] = tok;
        if (tok == p.eot) {
            done[b] = 1;
            out_len[b] = cur_len - n_initial;
            atomicAdd(&ctl->n_done, 1);
        } else {
            out_len[b] = cur_len - n_initial + 1;
        }
    }
}

__global__ void advance_kernel(DecodeCtl* ctl) {
    pdl_prologue();
    ctl->step += 1;
}

int launch_sample(const __half* logits, const unsigned char* suppress_mask, int* tokens, const int* /*initial_tokens*/,
                  float* sum_logprob, float* no_speech_prob, int* out_len, unsigned char* done, DecodeCtl* ctl, const DecodeParams& p,
                  cudaStream_t s) {
    const size_t smem = ((size_t)p.n_vocab * 2 + 15) & ~size_t(15);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
        if (e != cudaSuccess) return set_error("sample attr: %s", cudaGetErrorString(e));
        attr = true;
    }
    if (smem > 110 * 1024) return set_error("sample: vocab %d too large", p.n_vocab);
    launch_k(sample_kernel, dim3(p.B), dim3(kSampleThreads), smem, s, logits, suppress_mask, tokens, sum_logprob, no_speech_prob, out_len, done, ctl, p);
    WJB_CHECK_LAUNCH("sample");
    return 0;
}

int launch_advance(DecodeCtl* ctl, cudaStream_t s) {
    launch_k(advance_kernel, dim3(1), dim3(1), 0, s, ctl);
    WJB_CHECK_LAUNCH("advance");
    return 0;
}

int sample_init() {
    cudaError_t e = cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    if (e != cudaSuccess) return set_error("sample attr: %s", cudaGetErrorString(e));
    return 0;
}

}  // namespace wjb
