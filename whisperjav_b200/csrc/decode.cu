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

// parity_stride != 0: beam search keeps two token buffers `parity_stride` ints apart; position `step` is read from buffer step & 1
// lnstat (n_sites > 0): [n_sites][B][2] fixed-point (x 2^20) sum / sum of squares of the residual-stream rows at every LayerNorm
// site of the step (gemm_step.cu): this kernel stores its row's statistics at site 0 and zeroes the row's accumulators of the others
__global__ void embed_kernel(const int* __restrict__ tokens, int tokens_stride, long long parity_stride, const __half* __restrict__ emb,
                             const __half* __restrict__ pos, __half* __restrict__ x, const DecodeCtl* __restrict__ ctl, int n,
                             long long* __restrict__ lnstat, int n_sites) {
    pdl_prologue();
    const int b = blockIdx.x, B = gridDim.x;
    const int step = ctl->step;
    const int tok = tokens[(step & 1) * parity_stride + (long long)b * tokens_stride + step];
    const __half2* e = reinterpret_cast<const __half2*>(emb + (long long)tok * n);
    const __half2* p = reinterpret_cast<const __half2*>(pos + (long long)step * n);
    __half2* o = reinterpret_cast<__half2*>(x + (long long)b * n);
    long long sfx = 0, qfx = 0;
    for (int i = threadIdx.x; i < n / 2; i += blockDim.x) {
        const float2 a = __half22float2(e[i]), c = __half22float2(p[i]);
        const __half2 v = __floats2half2_rn(a.x + c.x, a.y + c.y);
        o[i] = v;
        const float2 f = __half22float2(v);
        sfx += __float2ll_rn(f.x * 1048576.0f) + __float2ll_rn(f.y * 1048576.0f);   // exact (see gemm_step.cu)
        qfx += __float2ll_rn(f.x * f.x * 1048576.0f) + __float2ll_rn(f.y * f.y * 1048576.0f);
    }
    if (n_sites > 0) {
        __shared__ long long red[2][4];
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) {
            sfx += __shfl_xor_sync(0xffffffffu, sfx, o2);
            qfx += __shfl_xor_sync(0xffffffffu, qfx, o2);
        }
        if ((threadIdx.x & 31) == 0) {
            red[0][threadIdx.x >> 5] = sfx;
            red[1][threadIdx.x >> 5] = qfx;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            lnstat[2 * b] = red[0][0] + red[0][1] + red[0][2] + red[0][3];
            lnstat[2 * b + 1] = red[1][0] + red[1][1] + red[1][2] + red[1][3];
        }
        for (int k = 1 + threadIdx.x; k < n_sites; k += blockDim.x) {
            lnstat[((long long)k * B + b) * 2] = 0;
            lnstat[((long long)k * B + b) * 2 + 1] = 0;
        }
    }
}

int launch_embed(const int* tokens, int tokens_stride, const __half* emb, const __half* pos, __half* x, const DecodeCtl* ctl, int B,
                 int n, cudaStream_t s, long long parity_stride, long long* lnstat, int n_sites) {
    launch_k(embed_kernel, dim3(B), dim3(128), 0, s, tokens, tokens_stride, parity_stride, emb, pos, x, ctl, n, lnstat, n_sites);
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

// ---- logit filters shared by the greedy and the beam kernels (decoding.py: SuppressBlank, SuppressTokens, ApplyTimestampRules)
struct RowState {
    int last_ts, pen_ts, have_ts, ts_last;
};
__device__ inline RowState row_state(const int* trow, int cur_len, int n_initial, const DecodeParams& p) {
    RowState st{0, 0, 0, 0};
    const int ns = cur_len - n_initial;  // sampled so far
    if (ns >= 0 && p.apply_timestamp_rules) {
        st.last_ts = ns >= 1 && trow[cur_len - 1] >= p.timestamp_begin;
        st.pen_ts = ns < 2 || trow[cur_len - 2] >= p.timestamp_begin;
        for (int i = cur_len - 1; i >= n_initial; --i) {
            if (trow[i] >= p.timestamp_begin) {
                st.have_ts = 1;
                st.ts_last = trow[i];
                break;
            }
        }
        if (st.have_ts && !(st.last_ts && !st.pen_ts)) st.ts_last += 1;
    }
    return st;
}
__device__ inline bool token_masked(int v, bool first, const RowState& st, const DecodeParams& p, const unsigned char* suppress_mask) {
    if (p.suppress_blank && first && (v == p.blank_token || v == p.eot)) return true;
    if (suppress_mask && suppress_mask[v]) return true;
    if (p.apply_timestamp_rules) {
        const int tsb = p.timestamp_begin;
        if (v == p.no_timestamps) return true;
        if (st.last_ts) {
            if (st.pen_ts) {
                if (v >= tsb) return true;
            } else {
                if (v < p.eot) return true;
            }
        }
        if (st.have_ts && v >= tsb && v < st.ts_last) return true;
        if (first) {
            if (v < tsb) return true;
            if (p.max_initial_timestamp_index >= 0 && v > tsb + p.max_initial_timestamp_index) return true;
        }
    }
    return false;
}

// ---- beam search step (decoding.py::BeamSearchDecoder.update), one CTA per audio window --------------------------------
// Rows a*beam .. a*beam+beam-1 are the live beams of window a.  Per row: filters, log-softmax, the beam+1 best tokens; then one
// thread ranks the beam*(beam+1) candidates (all beams are identical at the first sampled position: only beam 0 counts there,
// which is what upstream's dict of sequences amounts to), EOT candidates met on the way down join the window's finished list
// while it has room (max_candidates = round(beam * patience)), the best `beam` live ones become the next rows.  Instead of
// permuting the self-attention cache, every row carries the table of physical cache rows its history lives in (`anc`): the new
// row copies its parent's table and appends the parent's row for the current position.  Token rows, the tables and the scores
// are double buffered by step parity.
struct BeamCand {
    float score;
    int tok, src, order;
};

__global__ void __launch_bounds__(kSampleThreads)
beam_select_kernel(const __half* __restrict__ logits, const unsigned char* __restrict__ suppress_mask, const BeamBufs bb,
                   float* __restrict__ no_speech_prob, unsigned char* __restrict__ done, DecodeCtl* __restrict__ ctl, const DecodeParams p) {
    extern __shared__ __half srow[];
    __shared__ ArgMax sh_am[kSampleThreads / 32];
    __shared__ float sh_f[kSampleThreads / 32];
    __shared__ RowState sh_st;
    __shared__ BeamCand cand[kMaxBeam * (kMaxBeam + 1)];
    __shared__ int keep_src[kMaxBeam], keep_tok[kMaxBeam];
    __shared__ float keep_score[kMaxBeam];
    __shared__ int fin_src[kMaxBeam * (kMaxBeam + 1)], fin_slot[kMaxBeam * (kMaxBeam + 1)];
    __shared__ int n_fin_new, sh_complete;
    pdl_prologue();
    const int a = blockIdx.x, tid = threadIdx.x;
    const int beam = bb.beam;
    const int step = ctl->step;
    const int cur_len = step + 1;
    const int n_initial = ctl->n_initial;
    const bool at_sot = (step == ctl->sot_index);
    const bool sampling = cur_len >= n_initial;
    if (!at_sot && (!sampling || bb.audio_done[a])) return;
    const int V = p.n_vocab;
    const int par = step & 1;
    const int* tok_cur = bb.tokens + (long long)par * bb.tokens_parity_stride;
    int* tok_nxt = bb.tokens + (long long)(par ^ 1) * bb.tokens_parity_stride;
    const short* anc_cur = bb.anc + (long long)par * bb.anc_parity_stride;
    short* anc_nxt = bb.anc + (long long)(par ^ 1) * bb.anc_parity_stride;
    const float* slp_cur = bb.sum_logprob + par * bb.rows;
    float* slp_nxt = bb.sum_logprob + (par ^ 1) * bb.rows;
    const bool first = (cur_len == n_initial);
    const int split = p.apply_timestamp_rules ? p.timestamp_begin : V;

    for (int j = 0; j < beam; ++j) {
        const int row = a * beam + j;
        const __half* lrow = logits + (long long)row * p.logits_stride;
        __syncthreads();  // srow / sh_st of the previous beam are no longer read
        for (int v = tid; v < V; v += kSampleThreads) srow[v] = lrow[v];
        const int* trow = tok_cur + (long long)row * p.tokens_stride;
        if (tid == 0) sh_st = row_state(trow, cur_len, n_initial, p);
        __syncthreads();
        if (at_sot && j == 0) {
            ArgMax am{-INFINITY, 0};
            for (int v = tid; v < V; v += kSampleThreads) am = amax(am, ArgMax{__half2float(srow[v]), v});
            am = block_argmax(am, sh_am);
            float sm = 0.f;
            for (int v = tid; v < V; v += kSampleThreads) sm += __expf(__half2float(srow[v]) - am.v);
            sm = block_sum(sm, sh_f);
            if (tid == 0) no_speech_prob[a] = __expf(__half2float(srow[p.no_speech]) - am.v) / sm;
        }
        if (!sampling || bb.audio_done[a]) continue;
        if (first && j > 0) continue;  // identical beams: one set of candidates
        const RowState st = sh_st;
        ArgMax at{-INFINITY, 0x7fffffff}, as{-INFINITY, 0x7fffffff};
        for (int v = tid; v < V; v += kSampleThreads) {
            if (token_masked(v, first, st, p, suppress_mask)) {
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
        if (p.apply_timestamp_rules && s_ts > 0.f) ts_wins = (logf(s_ts) + mx) > at.v;
        if (ts_wins) {
            for (int v = tid; v < split; v += kSampleThreads) srow[v] = __float2half(-INFINITY);
        }
        const float lse = ts_wins ? mx + logf(s_ts) : mx + logf(s_text + s_ts);
        __syncthreads();
        // the beam + 1 best tokens of this row (logprobs.topk(beam + 1)): repeated block arg-max, ties to the lower id
        for (int r = 0; r <= beam; ++r) {
            ArgMax best{-INFINITY, 0x7fffffff};
            for (int v = tid; v < V; v += kSampleThreads) best = amax(best, ArgMax{__half2float(srow[v]), v});
            best = block_argmax(best, sh_am);
            if (tid == 0) {
                BeamCand c;
                c.score = (best.v == -INFINITY) ? -INFINITY : slp_cur[row] + (best.v - lse);
                c.tok = best.i;
                c.src = row;
                c.order = j * (kMaxBeam + 1) + r;
                cand[j * (beam + 1) + r] = c;
                if (best.i != 0x7fffffff) srow[best.i] = __float2half(-INFINITY);
            }
            __syncthreads();
        }
    }
    if (!sampling || bb.audio_done[a]) return;
    __syncthreads();
    if (tid == 0) {
        const int n_cand = (first ? 1 : beam) * (beam + 1);
        // insertion sort, descending score; equal scores keep upstream's insertion order (beam, then top-k rank)
        for (int i = 1; i < n_cand; ++i) {
            BeamCand c = cand[i];
            int k = i - 1;
            while (k >= 0 && (cand[k].score < c.score || (cand[k].score == c.score && cand[k].order > c.order))) {
                cand[k + 1] = cand[k];
                --k;
            }
            cand[k + 1] = c;
        }
        int saved = 0, nf = 0, fin_count = bb.fin_count[a];
        for (int i = 0; i < n_cand && saved < beam; ++i) {
            if (cand[i].score == -INFINITY) break;
            if (cand[i].tok == p.eot) {
                if (fin_count < bb.max_candidates) {  // the candidate list still has room
                    fin_src[nf] = cand[i].src;
                    fin_slot[nf] = fin_count;
                    bb.fin_score[a * bb.max_candidates + fin_count] = cand[i].score;
                    bb.fin_len[a * bb.max_candidates + fin_count] = cur_len + 1;
                    ++fin_count;
                    ++nf;
                }
            } else {
                keep_src[saved] = cand[i].src;
                keep_tok[saved] = cand[i].tok;
                keep_score[saved] = cand[i].score;
                ++saved;
            }
        }
        for (; saved < beam; ++saved) {  // cannot happen with beam + 1 candidates per row; keep the buffers defined anyway
            keep_src[saved] = keep_src[0];
            keep_tok[saved] = p.eot;
            keep_score[saved] = -INFINITY;
        }
        n_fin_new = nf;
        bb.fin_count[a] = fin_count;
        sh_complete = fin_count >= bb.max_candidates;
    }
    __syncthreads();
    // finished sequences: parent prefix + EOT
    for (int f = 0; f < n_fin_new; ++f) {
        const int* src = tok_cur + (long long)fin_src[f] * p.tokens_stride;
        int* dst = bb.fin_tokens + ((long long)a * bb.max_candidates + fin_slot[f]) * p.tokens_stride;
        for (int t = tid; t < cur_len; t += kSampleThreads) dst[t] = src[t];
        if (tid == 0 && cur_len < p.tokens_stride) dst[cur_len] = p.eot;
    }
    // next rows: parent's tokens / ancestry + the new token
    for (int i = 0; i < beam; ++i) {
        const int nrow = a * beam + i, src_row = keep_src[i];
        const int* src = tok_cur + (long long)src_row * p.tokens_stride;
        int* dst = tok_nxt + (long long)nrow * p.tokens_stride;
        const short* asrc = anc_cur + (long long)src_row * p.n_ctx;
        short* adst = anc_nxt + (long long)nrow * p.n_ctx;
        for (int t = tid; t < cur_len; t += kSampleThreads) {
            dst[t] = src[t];
            adst[t] = (t < step) ? asrc[t] : (short)src_row;  // position `step` was computed (and cached) by the parent row
        }
        if (tid == 0) {
            if (cur_len < p.tokens_stride) dst[cur_len] = keep_tok[i];
            slp_nxt[nrow] = keep_score[i];
        }
    }
    if (tid == 0 && sh_complete) {
        bb.audio_done[a] = 1;
        for (int i = 0; i < beam; ++i) done[a * beam + i] = 1;
        atomicAdd(&ctl->n_done, beam);
    }
}

int launch_beam_select(const __half* logits, const unsigned char* suppress_mask, const BeamBufs& bb, float* no_speech_prob,
                       unsigned char* done, DecodeCtl* ctl, const DecodeParams& p, cudaStream_t s) {
    if (bb.beam < 1 || bb.beam > kMaxBeam) return set_error("beam search: beam_size %d outside 1..%d", bb.beam, kMaxBeam);
    if (p.n_ctx > 32767) return set_error("beam search: n_ctx too large for the ancestry table");
    const size_t smem = ((size_t)p.n_vocab * 2 + 15) & ~size_t(15);
    if (smem > 110 * 1024) return set_error("beam search: vocab %d too large", p.n_vocab);
    launch_k(beam_select_kernel, dim3(bb.n_audio), dim3(kSampleThreads), smem, s, logits, suppress_mask, bb, no_speech_prob, done, ctl, p);
    WJB_CHECK_LAUNCH("beam_select");
    return 0;
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
    if (p.trace_logits) {
        __half* dst = p.trace_logits + ((long long)step * p.B + b) * p.logits_stride;
        for (int v = tid; v < V; v += kSampleThreads) dst[v] = row[v];
    }
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
    if (p.align_len) {
        // alignment pass: nothing is sampled.  Feed the forced token of the next position, record its probability under the raw
        // logits restricted to ids < eot (timing.py: logits[:, :eot].softmax), finish the row after its last token was fed.
        if (cur_len >= p.align_len[b]) {
            if (tid == 0) {
                done[b] = 1;
                out_len[b] = cur_len - n_initial;
                atomicAdd(&ctl->n_done, 1);
            }
            return;
        }
        const int tok = p.trace_forced[(long long)b * p.tokens_stride + cur_len];
        ArgMax a{-INFINITY, 0};
        for (int v = tid; v < p.eot; v += kSampleThreads) a = amax(a, ArgMax{__half2float(srow[v]), v});
        a = block_argmax(a, sh_am);
        float s = 0.f;
        for (int v = tid; v < p.eot; v += kSampleThreads) s += expf(__half2float(srow[v]) - a.v);
        s = block_sum(s, sh_f);
        if (tid == 0) {
            if (p.align_prob) p.align_prob[(long long)b * p.tokens_stride + cur_len] = tok < p.eot ? expf(__half2float(srow[tok]) - a.v) / s : 0.f;
            if (cur_len < p.tokens_stride) trow[cur_len] = tok;
            out_len[b] = cur_len - n_initial + 1;
        }
        return;
    }

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
        int tok = pick.i;
        if (p.trace_sampled && cur_len < p.tokens_stride) p.trace_sampled[(long long)b * p.tokens_stride + cur_len] = tok;
        if (p.trace_forced && cur_len < p.tokens_stride) tok = p.trace_forced[(long long)b * p.tokens_stride + cur_len];
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

int launch_sample(const __half* logits, const unsigned char* suppress_mask, int* tokens, float* sum_logprob, float* no_speech_prob, int* out_len, unsigned char* done, DecodeCtl* ctl, const DecodeParams& p,
                  cudaStream_t s) {
    const size_t smem = ((size_t)p.n_vocab * 2 + 15) & ~size_t(15);
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
    if (e == cudaSuccess) e = cudaFuncSetAttribute(beam_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    if (e != cudaSuccess) return set_error("sample attr: %s", cudaGetErrorString(e));
    return 0;
}

}  // namespace wjb
