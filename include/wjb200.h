/* libwjb200.so -- C-ABI of the B200-native WhisperJAV ASR hot path.
 *
 * The reference (meizhong986/WhisperJAV @ f7862f7) has no FFI of its own: its hot path is Python
 * calling third-party packages.  Every entry point below therefore cites the *reference call site*
 * whose arithmetic it replaces (paths relative to the reference repo root) and, in brackets, the
 * upstream function that call site reaches (openai-whisper 20250625 @ c0d2f62).  The Python side
 * of the boundary (whisperjav_b200/) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions: all pointers are device pointers unless marked host; `stream` is a cudaStream_t
 * passed as void*; no function allocates device memory (callers supply workspaces sized by the
 * *_workspace_bytes functions); every function returns 0 on success and a non-zero status on error,
 * with a message available from wjb_last_error(); nothing throws across the ABI.
 */
#ifndef WJB200_H_
#define WJB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WJB_ABI_VERSION 1

/* Whisper ModelDimensions [whisper/model.py::ModelDimensions]; the model name that selects them
 * arrives at whisperjav/modules/whisper_pro_asr.py:45,182 (`whisper.load_model(model_name)`). */
typedef struct wjb_dims {
    int32_t n_mels, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer;
    int32_t n_vocab, n_text_ctx, n_text_state, n_text_head, n_text_layer;
} wjb_dims;

typedef struct wjb_model wjb_model; /* opaque */

int wjb_abi_version(void);
const char* wjb_last_error(void); /* host string, thread-local */

/* ---- weights: one packed device blob --------------------------------------------------------
 * Replaces `whisper.load_model(name, device)` (whisper_pro_asr.py:182) as far as placement goes:
 * the caller converts a state_dict into this layout (whisperjav_b200/weights.py). */
size_t wjb_weights_bytes(const wjb_dims* dims);
int wjb_weight_count(const wjb_dims* dims);
/* index -> name (copied into name_buf), byte offset, byte size, dtype (0 = fp16, 1 = fp32) */
int wjb_weight_info(const wjb_dims* dims, int index, char* name_buf, int name_buf_len, size_t* offset, size_t* nbytes,
                    int* dtype);
int wjb_model_create(const wjb_dims* dims, const void* weights_blob, wjb_model** out);
void wjb_model_destroy(wjb_model* m);

/* ---- log-mel ------------------------------------------------------------------------------
 * Replaces `log_mel_spectrogram(audio, n_mels, padding=N_SAMPLES)` + `pad_or_trim` inside
 * `whisper_model.transcribe` (whisper_pro_asr.py:433) [whisper/audio.py], and
 * `WhisperProcessor(audio)` at modules/subtitle_pipeline/generators/anime_whisper.py:256-263 and
 * `WhisperFeatureExtractor` at modules/speech_segmentation/backends/whisperseg.py:376-380.
 *   audio      fp32 [n_clips][audio_stride]; n_samples (device int32 [n_clips]) valid samples each
 *   filters    fp32 [n_mels][201] Slaney mel filterbank
 *   out        fp16; time_major=1: [n_clips][out_rows][n_mels] with frame t at row row0+t
 *                    time_major=0: [n_clips][n_mels][n_frames]        (upstream layout)
 *   frames t >= n_samples/160 are literal zeros (pad_or_trim semantics); reflect_total = 0 means the
 *   signal is followed by >= 200 zeros (transcribe's 30 s padding), else the padded length at which the
 *   right edge reflects (480000 for the HF feature extractor).
 *   workspace: wjb_logmel_workspace_bytes(n_clips, n_mels) bytes. */
size_t wjb_logmel_workspace_bytes(int n_clips, int n_mels);
int wjb_logmel_f16(const float* audio, int64_t audio_stride, const int32_t* n_samples, int n_clips, int n_mels,
                   const float* filters, void* out, int time_major, int64_t out_clip_stride, int row0, int n_frames,
                   int reflect_total, void* workspace, void* stream);

/* ---- encoder ------------------------------------------------------------------------------
 * Replaces `model.encoder(mel)` reached from whisper_pro_asr.py:433 [whisper/model.py::AudioEncoder.forward]
 * and `model.generate()`'s encoder pass at generators/anime_whisper.py:279.
 *   mel_tm  fp16 [B][2*n_audio_ctx + 2][n_mels], time-major, rows 0 and 2*n_audio_ctx+1 zero (conv padding)
 *   out     fp16 [B][n_audio_ctx][n_audio_state] */
size_t wjb_encoder_workspace_bytes(const wjb_model* m, int batch);
int wjb_encoder_forward(wjb_model* m, const void* mel_tm, int batch, void* out, void* workspace, size_t workspace_bytes,
                        void* stream);

/* Parity-test hook: while `out` is non-NULL every following wjb_encoder_forward also copies the residual stream after blocks
 * every, 2*every, ... (1-based) to out[k][B][n_audio_ctx][n_audio_state] fp16 (k = block / every - 1); NULL = off. */
int wjb_encoder_set_tap(wjb_model* m, void* out, int every);

/* ---- cross-attention K/V projection (once per window) ---------------------------------------
 * Replaces the first-step `cross_attn.key/value(xa)` Linear calls cached by the KV hooks
 * [whisper/model.py::MultiHeadAttention.forward, whisper/decoding.py::PyTorchInference].
 *   kv_out fp16 [n_text_layer][B][2*n_text_head][n_audio_ctx][64]  (K heads, then V heads) */
size_t wjb_cross_kv_bytes(const wjb_model* m, int batch);
int wjb_cross_kv(wjb_model* m, const void* enc_out, int batch, void* kv_out, void* stream);

/* ---- greedy decode (persistent on device, one host read-back per run) ------------------------
 * Replaces `DecodingTask.run` / `_main_loop` [whisper/decoding.py] reached through
 * `decode_with_fallback` in `whisper_model.transcribe` (whisper_pro_asr.py:433) and HF
 * `model.generate(do_sample=False, num_beams=1)` (generators/anime_whisper.py:269-279). */
typedef struct wjb_decode_opts {
    int32_t n_initial;     /* forced prefix length (sot sequence [+ prompt]) = sample_begin */
    int32_t sot_index;     /* position of <|startoftranscript|> in the prefix */
    int32_t sample_len;    /* max sampled tokens (n_text_ctx / 2) */
    int32_t eot, no_speech, no_timestamps, timestamp_begin;
    int32_t suppress_blank, blank_token;
    int32_t apply_timestamp_rules;        /* 0 when without_timestamps */
    int32_t max_initial_timestamp_index;  /* -1 = none */
    int32_t tokens_stride;                /* ints per row of `tokens` (>= n_initial + sample_len + 1) */
    int32_t check_every;                  /* host polls the done counter every this many steps (0 = 8) */
    float temperature;                    /* 0 = greedy argmax; > 0 = Categorical(logits / T) by Gumbel-max, counter-based RNG */
    uint32_t seed;                        /* RNG stream for temperature > 0 (row, step and token id are mixed in) */
} wjb_decode_opts;

size_t wjb_decode_workspace_bytes(const wjb_model* m, int batch);
/*   cross_kv       from wjb_cross_kv
 *   suppress_mask  uint8 [n_vocab] (1 = suppressed) or NULL         [SuppressTokens]
 *   tokens         int32 [B][tokens_stride]; caller fills the first n_initial entries of every row
 *   sum_logprob    fp32 [B] (out), no_speech_prob fp32 [B] (out), out_len int32 [B] (out: sampled tokens before EOT)
 *   steps_run      host int (out, may be NULL): decoder steps executed */
int wjb_decode_greedy(wjb_model* m, const void* cross_kv, int batch, const wjb_decode_opts* opts,
                      const uint8_t* suppress_mask, int32_t* tokens, float* sum_logprob, float* no_speech_prob,
                      int32_t* out_len, void* workspace, size_t workspace_bytes, int* steps_run, void* stream);

/* Parity-test hook for wjb_decode_greedy (all NULL = off, the production state).  While set, every following greedy run
 *   - copies the raw fp16 logits of each step to logits_out[step][row][wjb_decode_logits_stride(m)] (upstream: the
 *     `logits = self.inference.logits(tokens, audio_features)` line of DecodingTask._main_loop, before the logit filters);
 *     logits_out_bytes must cover (n_initial - 1 + sample_len) steps;
 *   - records the id the device picked at position p of row b in sampled_out[b][tokens_stride] (p >= n_initial);
 *   - feeds forced_tokens[b][p] instead of the picked id (teacher forcing; same layout as `tokens`).
 * The step graph is re-captured when the hook changes; nothing is read from these pointers after a run returns. */
int wjb_decode_logits_stride(const wjb_model* m);
int wjb_decode_set_trace(wjb_model* m, void* logits_out, size_t logits_out_bytes, int32_t* sampled_out, const int32_t* forced_tokens);

/* ---- word-level timestamps (openai-whisper timing.py::find_alignment; `word_timestamps=True` in every reference preset,
 * config/components/asr/openai_whisper.py:213,247,281) -------------------------------------------------------------------
 * Step 1, the teacher-forced pass: with wjb_decode_set_align set, a wjb_decode_greedy run over tokens = [sot sequence,
 * <|notimestamps|>, text tokens, <|endoftext|>] (forced through wjb_decode_set_trace, n_initial = len(sot sequence) + 1,
 * sample_len = max n_tokens - n_initial + 1) feeds every row its n_tokens[b] tokens (no EOT latch) and records
 *   qk_out          fp16 [B][n_sel][n_steps][n_audio_ctx]: the scaled cross-attention scores q.k / sqrt(64) of every position, for
 *                   the alignment heads = every head of layers n_text_layer / 2 .. (upstream's default `alignment_heads`);
 *                   n_sel = (n_text_layer - n_text_layer / 2) * n_text_head, n_steps >= the run's step count;
 *   token_prob_out  fp32 [B][tokens_stride]: softmax(logits[: eot])[token fed at position p] (timing.py `text_token_probs`).
 * Step 2, wjb_align_dtw: softmax over the first n_frames2[b] frames, standardisation over the n_tokens[b] token rows, median
 * filter (medfilt_width 7, or 1 = off), mean over the heads -> matrix fp32 [B][n_steps][n_audio_ctx]; DTW over -matrix rows
 * row_begin[b] .. row_begin[b] + n_rows[b] - 1 -> jump_frames int32 [B][n_steps]: for DTW row i the first frame of its path
 * segment (timing.py `time_indices[jumps]`).  All int32 arrays are device memory.  Word grouping and the segment adjustments
 * are host logic (whisperjav_b200/timing.py). */
size_t wjb_align_qk_bytes(const wjb_model* m, int batch, int n_steps);
int wjb_decode_set_align(wjb_model* m, void* qk_out, int n_steps, const int32_t* n_tokens, float* token_prob_out);
/* Step 1 as ONE pass over whole sequences (every position of every window a GEMM row: one sweep of the decoder weights instead
 * of one decode step per position): tokens int32 [B][tokens_stride] holds the n_tokens[b] tokens of every row (device memory),
 * n_steps (a multiple of 8, <= 256) >= max n_tokens is the padded row count per window; same qk_out / token_prob_out layout as
 * above (probabilities are softmax(logits[: eot])); cross_kv from wjb_cross_kv.  Workspace: wjb_align_prefill_workspace_bytes(m, batch, n_steps). */
size_t wjb_align_prefill_workspace_bytes(const wjb_model* m, int batch, int n_steps);
int wjb_align_prefill(wjb_model* m, const void* cross_kv, const int32_t* tokens, int tokens_stride, const int32_t* n_tokens, int batch,
                      int n_steps, int eot, void* qk_out, float* token_prob_out, void* workspace, size_t workspace_bytes, void* stream);
size_t wjb_align_workspace_bytes(const wjb_model* m, int batch, int n_steps);
int wjb_align_dtw(wjb_model* m, const void* qk, int batch, int n_steps, const int32_t* n_tokens, const int32_t* row_begin,
                  const int32_t* n_rows, const int32_t* n_frames2, int medfilt_width, float* matrix, int32_t* jump_frames,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ---- beam search decode (openai-whisper decoding.py::BeamSearchDecoder at temperature 0; replaces the decode inside
 * whisper_pro_asr.py:433 when the preset sets beam_size, config/components/asr/openai_whisper.py:225-292) -----------------
 * Rows are windows x beams (row = window * beam_size + beam).  The self-attention cache is never permuted: every row carries the
 * table of physical cache rows its history lives in (`anc`).  All buffers are device memory owned by the caller:
 *   tokens       int32 [2][rows][tokens_stride]  both copies pre-filled with the n_initial prompt tokens of every row
 *   anc          int16 [2][rows][n_text_ctx]     both copies pre-filled with the row index
 *   sum_logprob  fp32  [2][rows]                 zeroed; after the run the live beams' scores are in copy (steps_run & 1)
 *   fin_tokens   int32 [n_audio][max_candidates][tokens_stride], fin_score fp32, fin_len int32 (prompt + sampled + EOT),
 *   fin_count    int32 [n_audio] zeroed, audio_done uint8 [n_audio] zeroed
 * After the run the live token rows are in copy (steps_run & 1); BeamSearchDecoder.finalize and the MaximumLikelihoodRanker are
 * host logic on these buffers (whisperjav_b200/model.py).  Workspace: wjb_decode_workspace_bytes(m, rows). */
typedef struct wjb_beam_bufs {
    int32_t n_audio, beam_size, max_candidates; /* max_candidates = round(beam_size * patience) */
    int32_t* tokens;
    int16_t* anc;
    float* sum_logprob;
    int32_t* fin_tokens;
    float* fin_score;
    int32_t* fin_len;
    int32_t* fin_count;
    uint8_t* audio_done;
} wjb_beam_bufs;
int wjb_decode_beam(wjb_model* m, const void* cross_kv, const wjb_beam_bufs* bufs, const wjb_decode_opts* opts,
                    const uint8_t* suppress_mask, float* no_speech_prob, void* workspace, size_t workspace_bytes, int* steps_run,
                    void* stream);

/* ---- building blocks exposed for parity tests and profiling ---------------------------------- */
/* out[r][n] = epilogue(sum_k A[r][k] W[n][k]); flags: 1 = GELU (exact erf).  All fp16, fp32 accumulate. */
int wjb_gemm_f16(const void* A, int64_t a_row_stride, int64_t a_batch_stride, int rows_per_batch, int n_batch, int K,
                 const void* W, int N, int ldw, const void* bias, const void* residual, void* out, int64_t out_row_stride,
                 int64_t out_batch_stride, int flags, int block_n, void* stream);
/* the same GEMM for few rows (<= 128, one decoder step) with K sliced over `splits` CTAs per output tile (-1 = as many as fill
 * the SMs); the slices meet in `workspace` (fp32) and the last CTA to arrive adds them in slice order, so results do not
 * depend on timing.  workspace: wjb_gemm_splitk_workspace_bytes() bytes, zeroed once by the caller, not shared with a
 * GEMM running concurrently. */
size_t wjb_gemm_splitk_workspace_bytes(void);
int wjb_gemm_f16_splitk(const void* A, int64_t a_row_stride, int rows, int K, const void* W, int N, int ldw, const void* bias,
                        const void* residual, void* out, int64_t out_row_stride, int flags, int block_n, int splits,
                        void* workspace, size_t workspace_bytes, void* stream);
/* the decoder-step GEMM proper (rows <= 128, N % 64 == 0, K >= 512): a thread-block cluster of `cluster` CTAs (1|2|4|8,
 * 0 = auto) slices K of one 128 x block_n (64|128|256, 0 = auto) output tile and the partial tiles meet in distributed shared
 * memory, added in rank order.  w_constant = 1 promises W is not written by earlier work on the stream (model weights): its
 * loads are then issued ahead of the programmatic-dependent-launch wait. */
int wjb_gemm_step_f16(const void* A, int64_t a_row_stride, int rows, int K, const void* W, int N, int ldw, const void* bias,
                      const void* residual, void* out, int64_t out_row_stride, int flags, int block_n, int cluster,
                      int w_constant, void* stream);
/* the same with the LayerNorm in front of the Linear fused in (ln_gamma / ln_beta fp16 [K], eps 1e-5, fp32 statistics over the K
 * columns of every row of A; both NULL = plain): the cluster exchanges per-row partial sums through distributed shared memory */
int wjb_gemm_step_ln_f16(const void* A, int64_t a_row_stride, int rows, int K, const void* ln_gamma, const void* ln_beta,
                         const void* W, int N, int ldw, const void* bias, const void* residual, void* out, int64_t out_row_stride,
                         int flags, int block_n, int cluster, int w_constant, void* stream);
/* the decode step's form of it: no exchange.  ln_stats int64 [rows][2] = fixed-point (x 2^20) sum and sum of squares of every row of
 * A, accumulated by the launches that produced A: a launch given out_stats (int64 [rows][2], zeroed by the caller) ADDS the same
 * statistics of the fp16 values it stores, with 64-bit integer atomics (order-independent, bit-reproducible).  ln_stats NULL with
 * gamma / beta set = the exchanging form above; out_stats NULL = none.  [whisper/model.py::ResidualAttentionBlock: attn_ln,
 * cross_attn_ln, mlp_ln in front of the query / fc1 Linears, reached from whisper_pro_asr.py:433] */
int wjb_gemm_step_stats_f16(const void* A, int64_t a_row_stride, int rows, int K, const void* ln_gamma, const void* ln_beta,
                            const int64_t* ln_stats, const void* W, int N, int ldw, const void* bias, const void* residual, void* out,
                            int64_t out_row_stride, int64_t* out_stats, int flags, int block_n, int cluster, int w_constant,
                            void* stream);
/* debugging aid: CTA 0 of every following GEMM launch writes a timeline (SM clock, global timer per pipeline event) into
 * `buf` (device, (32 + 512 * 32) uint64, zeroed by the caller; a ring of the last 512 launches); NULL switches it off. */
void wjb_debug_gemm_trace(void* buf);
/* debugging aid: launch the building blocks below with the programmatic-dependent-launch attribute, as the decode graph does */
void wjb_debug_set_pdl(int on);
/* debugging / measurement aid: placement options of the decode step (bit 0 LayerNorms folded into the step GEMMs through row
 * statistics, bit 1 L2 prefetch of the next Linear's weights, bit 2 cross-attention K/V primed ahead of the dependency wait, bit 3
 * evict-first K/V loads).  Results do not depend on bits 1-3; bit 0 moves rounding by < 1 fp16 ulp of a LayerNorm output in rare
 * elements.  The library ships with bits 2 and 3 on (what measured fastest, profiles/r2_decode_flags.json); scripts/decode_flags_probe.py
 * measures the combinations. */
void wjb_debug_set_decode_flags(int flags);
int wjb_debug_get_decode_flags(void);
int wjb_layernorm_f16(const void* x, const void* gamma, const void* beta, void* out, int rows, int n, void* stream);
int wjb_attention_encoder_f16(const void* qkv, void* out, int batch, int T, int n_head, void* stream);
/* single decoder step pieces */
/* self-attention at *position (device int32): appends the k, v of qkv [B][3 * n_state] to kv_cache [B][2 * n_head][n_ctx][64]
 * and attends over positions 0..*position */
int wjb_attention_self_f16(const void* qkv, void* kv_cache, void* out, const int32_t* position, int batch, int n_head, int n_ctx,
                           void* stream);
int wjb_attention_cross_f16(const void* q, const void* kv, void* out, int batch, int n_head, int T, void* stream);
/* beam search: rows = windows * beams queries, row r attends over the K/V of window r / beams (kv [windows][2 * n_head][T][64]); for
 * beams <= 4 the beams of a window share one pass over its K/V (upstream repeats the audio features per beam,
 * whisper/decoding.py::DecodingTask.run `audio_features.repeat_interleave(self.n_group, dim=0)`) */
int wjb_attention_cross_beam_f16(const void* q, const void* kv, void* out, int rows, int n_head, int T, int beams, void* stream);

/* ---- per-kernel-class timing of wjb_encoder_forward (CUDA events on the launching stream) ------
 * classes: 0 = tcgen05 GEMM (conv1/conv2/linear), 1 = encoder attention, 2 = LayerNorm.  Off by default. */
void wjb_profile_enable(int on);
int wjb_profile_read(float* ms_by_class, int* launches_by_class, int n_classes);

/* ---- voice-activity gate ----------------------------------------------------------------------
 * Replaces the per-window model calls inside `get_speech_timestamps(tensor, jit_model)` at
 * modules/speech_segmentation/backends/silero.py:269-273 (and silero_v6.py:205-210; the per-hop
 * `TenVad.process` loop at backends/ten.py:232-239): audio -> per-window speech probability.
 * Architecture and weight layout: see whisperjav_b200/vad.py.  probs fp32 [n_clips][n_windows]. */
size_t wjb_vad_weights_bytes(void);
size_t wjb_vad_workspace_bytes(int n_clips, int n_windows);
int wjb_vad_forward(const float* audio, int64_t audio_stride, const int32_t* n_samples, int n_clips, const void* weights,
                    float* probs, int n_windows, void* workspace, void* stream);

/* Frame head of the WhisperSeg-class gate (speech_segmentation/backends/whisperseg.py:355-393: 30 s window -> 80-mel ->
 * Whisper-base-shaped encoder -> one logit per 20 ms frame -> sigmoid): prob[r] = sigmoid(x[r] . w + bias) for `rows` encoder
 * frames x fp16 [rows][n]; the encoder itself is wjb_encoder_forward on a base-sized model handle. */
int wjb_frame_head_f16(const void* x, const void* w, float bias, float* prob, int rows, int n, void* stream);

/* ---- scene-detection energy gate ------------------------------------------------------------------
 * Replaces the per-block energy computation inside `auditok.split(audio_bytes, ..., energy_threshold=...)` at
 * modules/scene_detection_backends/auditok_backend.py:392 (pass 1, the whole stream) and :567 (pass 2, every oversized
 * chapter) [auditok 0.3.0 `signal.calculate_energy` on 50 ms blocks of the int16 bytes built by
 * `(audio * 32767).astype(np.int16)` at :379 / :555].
 *   audio         fp32 [n_audio] mono samples of the stream
 *   region_start / region_len  int64 [n_regions]: sample range of every region of this pass (pass 1: one region = the stream)
 *   window_base   int64 [n_regions + 1]: windows before region r (window_base[n_regions] = n_windows); region r has
 *                 ceil(region_len[r] / window) windows, the last one may be short
 *   sumsq         uint64 [n_windows]: EXACT sum over the window of (int16 sample)^2 -- the host derives upstream's float64 dB
 *                 value and the `>= energy_threshold` flag from it bit-identically (whisperjav_b200/scenes.py). */
int wjb_scene_energy(const float* audio, int64_t n_audio, const int64_t* region_start, const int64_t* region_len,
                     const int64_t* window_base, int n_regions, int window, uint64_t* sumsq, int64_t n_windows, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WJB200_H_ */
