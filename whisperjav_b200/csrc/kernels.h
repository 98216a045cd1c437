// Internal launch interfaces shared by the .cu files of libwjb200.so (not part of the C-ABI).
#pragma once
#include "common.cuh"

namespace wjb {

// ---- error plumbing (api.cu) -------------------------------------------------------------
int set_error(const char* fmt, ...);  // records the message, returns a non-zero status
int sm_count();
int ensure_init();  // per-device kernel attribute set-up (api.cu)
typedef CUresult (*tensor_map_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
tensor_map_encode_fn get_tensor_map_encoder();

#define WJB_CHECK_LAUNCH(name)                                                              \
    do {                                                                                    \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess) return ::wjb::set_error("%s: %s", name, cudaGetErrorString(e__)); \
    } while (0)

// ---- launches: optional programmatic-dependent-launch attribute (set around the decoder step) -------------
bool pdl_enabled();
void set_pdl(bool on);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    int n = 0;
    if (pdl_enabled()) {
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        n = 1;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- GEMM (gemm_tc.cu) ---------------------------------------------------------------------
enum { GEMM_GELU = 1, GEMM_HEADSPLIT = 2 };
struct GemmArgs {
    const __half* A = nullptr;      // activations, addressed as [n_batch][rows_per_batch][K] with the strides below
    long long a_row_stride = 0;     // halfs (may be < K: overlapping im2col rows)
    long long a_batch_stride = 0;   // halfs
    int rows_per_batch = 0, n_batch = 1, K = 0;
    const __half* W = nullptr;      // [N][ldw]
    int N = 0, ldw = 0;
    const __half* bias = nullptr;
    const __half* residual = nullptr;
    const float* pos = nullptr;
    __half* out = nullptr;
    long long out_row_stride = 0, out_batch_stride = 0;
    int flags = 0;
    int hs_T = 0, hs_H = 0;
    int block_n = 0;                // 0 = auto
    // split-K (few-row GEMMs that cannot fill the SMs with output tiles alone): 0/1 = off, > 1 = that many k slices,
    // -1 = as many as fill the SMs. Needs a workspace that no concurrently running GEMM shares.
    int splits = 0;
    float* splitk_ws = nullptr;     // splitk_ws_bytes
    size_t splitk_ws_bytes = 0;
    unsigned* splitk_cnt = nullptr; // kSplitKCounters zeroed counters
};
constexpr int kSplitKCounters = 1024;
int launch_gemm(const GemmArgs& a, cudaStream_t stream);
void gemm_set_trace(void* buf);  // debugging aid: device buffer of (32 + 512 * 32) u64 receiving CTA-0 timelines, null = off
int gemm_init();  // set kernel attributes up front (outside any stream capture)
// decoder-step variant (gemm_step.cu): <= 128 rows, a cluster of `cluster` CTAs slices K of one output tile and meets in
// distributed shared memory; weights flagged constant are fetched ahead of the programmatic-dependent-launch wait
struct StepGemmArgs {
    const __half* A = nullptr;
    long long a_row_stride = 0;
    int rows = 0, K = 0;
    const __half* W = nullptr;
    int N = 0, ldw = 0;
    const __half* bias = nullptr;
    const __half* residual = nullptr;
    __half* out = nullptr;
    long long out_row_stride = 0;
    int flags = 0;      // GEMM_GELU
    int block_n = 0;    // 64 | 128 | 256, 0 = auto
    int cluster = 0;    // 1 | 2 | 4 | 8, 0 = auto
    bool w_constant = false;
    const __half* ln_gamma = nullptr;  // LayerNorm over the K columns of A applied inside the kernel (both or neither)
    const __half* ln_beta = nullptr;
    const long long* ln_stats = nullptr;  // with ln_gamma / ln_beta: [rows][2] fixed-point (x 2^20) sum and sum of squares of every row of A,
                                          // accumulated by the producers of A (out_stats of an earlier launch / the embedding kernel)
    long long* out_stats = nullptr;       // [rows][2]: this launch adds the same statistics of the rows it stores (zeroed by the caller)
    const void* prefetch = nullptr;       // constant bytes (the next Linear's weights) to pull into L2 while this launch runs
    size_t prefetch_bytes = 0;
};
bool gemm_step_supported(int rows, int N, int K);
int launch_gemm_step(const StepGemmArgs& a, cudaStream_t stream);
int gemm_step_init();
void gemm_step_set_trace(void* buf);
// ---- elementwise / normalisation (elementwise.cu) ------------------------------------------
int launch_layernorm(const __half* x, const __half* gamma, const __half* beta, __half* out, int rows, int n, cudaStream_t s);
int launch_frame_head(const __half* x, const __half* w, float bias, float* prob, int rows, int n, cudaStream_t s);
int launch_im2col_k3(const __half* xpad, __half* out, int B, int T_out, int C, int stride, int T_in_padded, cudaStream_t s);

// ---- log-mel (logmel.cu) -------------------------------------------------------------------
struct LogmelArgs {
    const float* audio;          // [n_clips][audio_stride] fp32
    long long audio_stride;
    const int* n_samples;        // device [n_clips]: valid samples per clip (rest is zero)
    int n_clips;
    int n_mels;
    const float* filters;        // device [n_mels][201] fp32 (Slaney mel filterbank)
    __half* out;                 // see layout
    int time_major;              // 1: out[clip][row0 + t][n_mels] (row stride n_mels); 0: out[clip][n_mels][n_frames]
    long long out_clip_stride;   // halfs
    int row0;                    // first written row in the time-major layout (1 = leave a zero pad row for conv1)
    int n_frames;                // frames written per clip; frames >= n_samples/160 are literal zeros (pad_or_trim)
    int reflect_total;           // 0: signal is followed by >= 200 zeros (transcribe's padding=N_SAMPLES); else reflect at this length
    float* clip_max;             // device [n_clips] workspace (log10 max per clip)
    int* mel_range;              // device [2*n_mels] workspace (non-zero span of each filter row)
};
int launch_logmel(const LogmelArgs& a, cudaStream_t s);

// ---- attention (attention.cu) --------------------------------------------------------------
// Encoder self-attention over qkv [B*T][3n] (q | k | v, heads of 64), out [B*T][n].
int launch_attn_encoder(const __half* qkv, __half* out, int B, int T, int H, cudaStream_t s);
int attn_init();
int attn_cross_init();
// Decoder single-token self-attention with HBM KV cache [B][2H][n_ctx][64] (K heads then V heads).
// anc (beam search, may be null): [2][B][n_ctx] physical cache row holding each position of a row's history, buffer = step & 1
int launch_attn_dec_self(const __half* qkv, __half* kv_cache, __half* out, const int* step, const unsigned char* done, int B, int H,
                         int n_ctx, cudaStream_t s, const short* anc = nullptr, long long anc_parity_stride = 0);
// Decoder single-token cross-attention over kv [B][2H][T][64].
// kv_div (beam search): row b reads the K/V of window b / kv_div
// capture (word-timestamp alignment pass): scaled scores of the position *step are also written, fp16, to
// base[b * b_stride + h * head_stride + *step * T + t]
struct CrossCapture {
    __half* base = nullptr;
    long long b_stride = 0, head_stride = 0;
    const int* step = nullptr;
};
struct CrossTuning {              // decode-step placement of the cross-attention stream (all off = round-1 behaviour)
    int early_kv = 0;             // prime the K/V ring before the programmatic-dependent-launch wait (K/V are constants of the run)
    int evict_first = 0;          // K/V bulk copies carry an L2 evict-first policy
    const void* pf_ptr = nullptr; // constant bytes (the following Linear's weights) the grid pulls into L2
    unsigned long long pf_bytes = 0;
};
int launch_attn_dec_cross(const __half* q, const __half* kv, __half* out, const unsigned char* done, int B, int H, int T,
                          cudaStream_t s, int kv_div = 1, const CrossCapture* capture = nullptr, const CrossTuning* tuning = nullptr);

// ---- decoder token logic (decode.cu) -------------------------------------------------------
struct DecodeCtl {            // device-resident control block, one per decode run
    int step;                 // absolute position of the token being fed (0-based)
    int n_initial;            // number of forced initial tokens (sample_begin)
    int sot_index;
    int max_steps;            // sample_len
    int n_done;
    float temperature;        // 0 = argmax; read by sample_kernel at run time so the step graph does not depend on it
    unsigned seed;
};
struct DecodeParams {
    int B, n_vocab, logits_stride;
    int eot, no_speech, no_timestamps, timestamp_begin;
    int suppress_blank, blank_token, apply_timestamp_rules, max_initial_timestamp_index;  // -1 = none
    int n_ctx;
    int tokens_stride;        // ints per row in tokens[]
    // test / diagnostic hook (all null in production): raw logits [step][B][logits_stride], the ids the kernel picked
    // [B][tokens_stride], ids to feed instead (teacher forcing) [B][tokens_stride]
    __half* trace_logits = nullptr;
    int* trace_sampled = nullptr;
    const int* trace_forced = nullptr;
    // word-timestamp alignment pass (teacher-forced, timing.py::find_alignment): row b is fed align_len[b] tokens in all (it is
    // finished after the step that feeds its last one, EOT included -- no EOT latch) and align_prob[b][p] receives
    // softmax(raw logits[: eot])[forced token at p]
    const int* align_len = nullptr;
    float* align_prob = nullptr;
};
int launch_embed(const int* tokens, int tokens_stride, const __half* emb, const __half* pos, __half* x, const DecodeCtl* ctl, int B,
                 int n, cudaStream_t s, long long parity_stride = 0, long long* lnstat = nullptr, int n_sites = 0);
// beam search state (decode.cu::beam_select_kernel); everything device memory owned by the caller
constexpr int kMaxBeam = 8;
struct BeamBufs {
    int n_audio, beam, rows, max_candidates;
    int* tokens;                 // [2][rows][tokens_stride], both buffers pre-filled with the initial tokens
    long long tokens_parity_stride;
    short* anc;                  // [2][rows][n_ctx] physical cache row of every position, both buffers pre-filled with the row id
    long long anc_parity_stride;
    float* sum_logprob;          // [2][rows], zeroed
    int* fin_tokens;             // [n_audio][max_candidates][tokens_stride]
    float* fin_score;            // [n_audio][max_candidates]
    int* fin_len;                // [n_audio][max_candidates] tokens incl. the initial ones and the closing EOT
    int* fin_count;              // [n_audio], zeroed
    unsigned char* audio_done;   // [n_audio], zeroed
};
int launch_beam_select(const __half* logits, const unsigned char* suppress_mask, const BeamBufs& bb, float* no_speech_prob,
                       unsigned char* done, DecodeCtl* ctl, const DecodeParams& p, cudaStream_t s);
int launch_sample(const __half* logits, const unsigned char* suppress_mask, int* tokens, float* sum_logprob,
                  float* no_speech_prob, int* out_len, unsigned char* done, DecodeCtl* ctl, const DecodeParams& p, cudaStream_t s);

int launch_advance(DecodeCtl* ctl, cudaStream_t s);

// ---- word-timestamp alignment (align.cu) ------------------------------------------------------
size_t align_workspace_bytes(int B, int n_sel, int n_steps, int T);
int launch_align(const __half* qk, int B, int n_sel, int n_steps, int T, const int* n_tok, const int* row_begin, const int* n_rows,
                 const int* n_frames2, int medfilt, float* matrix, int* jump, void* workspace, size_t ws_bytes, cudaStream_t s);

// ---- teacher-forced pass over whole sequences (prefill.cu) ---------------------------------------
int launch_prefill_embed(const int* tokens, int tok_stride, const int* n_tok, const __half* emb, const __half* pos, __half* x, int B, int Lp, int n,
                         cudaStream_t s);
int launch_prefill_self_attn(const __half* qkv, __half* out, const int* n_tok, int B, int Lp, int H, cudaStream_t s);
int launch_prefill_cross_attn(const __half* q, const __half* kv, __half* out, const int* n_tok, int B, int Lp, int H, int T, const CrossCapture* cap,
                              cudaStream_t s);
int launch_prefill_prob(const __half* logits, int stride, int r0, int rows, const int* tokens, int tok_stride, const int* n_tok, float* prob, int Lp,
                        int eot, cudaStream_t s);

// ---- VAD (vad.cu) --------------------------------------------------------------------------
struct VadArgs;
}  // namespace wjb
