// extern "C" surface of libwjb200.so (see include/wjb200.h) plus the host-side orchestration of the
// encoder forward, the cross-K/V projection and the CUDA-graph-replayed greedy decode loop.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/wjb200.h"
#include "kernels.h"

namespace wjb {

// ------------------------------------------------------------------ error plumbing
static thread_local char g_err[512] = "";

int set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

tensor_map_encode_fn get_tensor_map_encoder() {
    static tensor_map_encode_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<tensor_map_encode_fn>(p);
    }
    return fn;
}

int sample_init();

static bool g_pdl = false;
bool pdl_enabled() { return g_pdl; }
void set_pdl(bool on) { g_pdl = on; }

// ---- optional per-kernel-class timing (CUDA events on the launching stream) -----------------
enum { PC_GEMM = 0, PC_ATTN = 1, PC_LN = 2, PC_N = 3 };
struct ProfRec {
    int cls;
    cudaEvent_t a, b;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_ev_pool;
static cudaEvent_t prof_event() {
    if (!g_ev_pool.empty()) {
        cudaEvent_t e = g_ev_pool.back();
        g_ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
struct ProfScope {
    int idx = -1;
    cudaStream_t s;
    ProfScope(int cls, cudaStream_t st) : s(st) {
        if (!g_prof_on) return;
        ProfRec r{cls, prof_event(), prof_event()};
        cudaEventRecord(r.a, s);
        g_prof.push_back(r);
        idx = (int)g_prof.size() - 1;
    }
    ~ProfScope() {
        if (idx >= 0) cudaEventRecord(g_prof[idx].b, s);
    }
};

int logmel_init();
int vad_init();
int align_init();
int prefill_init();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: run the kernel set-up once for every device a caller uses
// (a process may hold one model per GPU).  Called at the top of every compute entry point, outside any stream capture.
int ensure_init() {
    static unsigned long long done_mask = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_error("no CUDA device");
    if (dev < 64 && (done_mask >> dev) & 1ull) return 0;
    if (int e = gemm_init()) return e;
    if (int e = gemm_step_init()) return e;
    if (int e = attn_init()) return e;
    if (int e = attn_cross_init()) return e;
    if (int e = sample_init()) return e;
    if (int e = logmel_init()) return e;
    if (int e = vad_init()) return e;
    if (int e = align_init()) return e;
    if (int e = prefill_init()) return e;
    if (dev < 64) done_mask |= 1ull << dev;
    return 0;
}

// ------------------------------------------------------------------ weight blob layout
struct Entry {
    std::string name;
    size_t offset, bytes;
    int dtype;  // 0 fp16, 1 fp32
};
struct Layout {
    std::vector<Entry> e;
    size_t total = 0;
    void add(const std::string& name, size_t elems, int dtype = 0) {
        const size_t bytes = elems * (dtype ? 4 : 2);
        e.push_back({name, total, bytes, dtype});
        total += (bytes + 255) & ~size_t(255);
    }
    size_t off(const std::string& name) const {
        for (auto& x : e)
            if (x.name == name) return x.offset;
        return (size_t)-1;
    }
};

static Layout build_layout(const wjb_dims& d) {
    Layout L;
    const size_t n = d.n_audio_state, C = d.n_mels;
    L.add("enc.conv1.w", n * 3 * C);
    L.add("enc.conv1.b", n);
    L.add("enc.conv2.w", n * 3 * n);
    L.add("enc.conv2.b", n);
    L.add("enc.pos", (size_t)d.n_audio_ctx * n, 1);
    for (int i = 0; i < d.n_audio_layer; ++i) {
        const std::string p = "enc." + std::to_string(i) + ".";
        L.add(p + "ln1.g", n);
        L.add(p + "ln1.b", n);
        L.add(p + "qkv.w", 3 * n * n);
        L.add(p + "qkv.b", 3 * n);
        L.add(p + "out.w", n * n);
        L.add(p + "out.b", n);
        L.add(p + "ln2.g", n);
        L.add(p + "ln2.b", n);
        L.add(p + "fc1.w", 4 * n * n);
        L.add(p + "fc1.b", 4 * n);
        L.add(p + "fc2.w", 4 * n * n);
        L.add(p + "fc2.b", n);
    }
    L.add("enc.ln_post.g", n);
    L.add("enc.ln_post.b", n);
    const size_t t = d.n_text_state;
    L.add("dec.emb", (size_t)d.n_vocab * t);
    L.add("dec.pos", (size_t)d.n_text_ctx * t);
    for (int i = 0; i < d.n_text_layer; ++i) {
        const std::string p = "dec." + std::to_string(i) + ".";
        L.add(p + "ln1.g", t);
        L.add(p + "ln1.b", t);
        L.add(p + "qkv.w", 3 * t * t);
        L.add(p + "qkv.b", 3 * t);
        L.add(p + "out.w", t * t);
        L.add(p + "out.b", t);
        L.add(p + "ln2.g", t);
        L.add(p + "ln2.b", t);
        L.add(p + "cq.w", t * t);
        L.add(p + "cq.b", t);
        L.add(p + "ckv.w", 2 * t * t);
        L.add(p + "ckv.b", 2 * t);
        L.add(p + "cout.w", t * t);
        L.add(p + "cout.b", t);
        L.add(p + "ln3.g", t);
        L.add(p + "ln3.b", t);
        L.add(p + "fc1.w", 4 * t * t);
        L.add(p + "fc1.b", 4 * t);
        L.add(p + "fc2.w", 4 * t * t);
        L.add(p + "fc2.b", t);
    }
    L.add("dec.ln.g", t);
    L.add("dec.ln.b", t);
    return L;
}

}  // namespace wjb

using namespace wjb;

struct wjb_model {
    wjb_dims d;
    const uint8_t* blob;
    Layout L;
    // decode graph cache
    cudaGraphExec_t graph = nullptr;
    const void* g_kv = nullptr;
    void* g_ws = nullptr;
    int g_B = 0;
    wjb_decode_opts g_opts;
    const void* g_mask = nullptr;
    int32_t* g_tokens = nullptr;
    float* g_slp = nullptr;
    float* g_nsp = nullptr;
    int32_t* g_len = nullptr;
    int* h_done = nullptr;  // pinned
    DecodeCtl* h_ctl = nullptr;  // pinned staging copy of the control block
    cudaStream_t own_stream = nullptr;  // decode runs here (the caller's stream may be the legacy stream, which cannot be captured)
    cudaEvent_t ev = nullptr;
    // test / diagnostic hook (wjb_decode_set_trace): raw logits of every step, sampled ids, teacher-forced ids
    __half* trace_logits = nullptr;
    size_t trace_logits_bytes = 0;
    int32_t* trace_sampled = nullptr;
    const int32_t* trace_forced = nullptr;
    const void *g_trace_logits = nullptr, *g_trace_sampled = nullptr, *g_trace_forced = nullptr;
    // word-timestamp alignment pass (wjb_decode_set_align): cross-attention score capture of the upper half of the layers
    __half* align_qk = nullptr;
    int align_steps = 0;
    const int32_t* align_len = nullptr;
    float* align_prob = nullptr;
    const void *g_align_qk = nullptr, *g_align_len = nullptr, *g_align_prob = nullptr;
    int g_align_steps = 0;
    int g_flags = -1;
    void* enc_tap = nullptr;  // test hook (wjb_encoder_set_tap): residual stream after block index k * enc_tap_every - 1
    int enc_tap_every = 0;
    const __half* h16(const std::string& name) const { return reinterpret_cast<const __half*>(blob + L.off(name)); }
    const float* f32(const std::string& name) const { return reinterpret_cast<const float*>(blob + L.off(name)); }
};

static bool dims_ok(const wjb_dims* d) {
    return d && d->n_audio_state % 64 == 0 && d->n_text_state % 64 == 0 && d->n_audio_state == d->n_audio_head * 64 &&
           d->n_text_state == d->n_text_head * 64 && d->n_mels % 8 == 0 && d->n_mels <= 128 && d->n_text_ctx <= 448 &&
           d->n_audio_ctx <= 1536 && d->n_audio_state <= 2048 && d->n_text_state <= 2048;
}

extern "C" {

int wjb_abi_version(void) { return WJB_ABI_VERSION; }
const char* wjb_last_error(void) { return g_err; }

size_t wjb_weights_bytes(const wjb_dims* dims) {
    if (!dims_ok(dims)) return 0;
    return build_layout(*dims).total;
}
int wjb_weight_count(const wjb_dims* dims) {
    if (!dims_ok(dims)) return 0;
    return (int)build_layout(*dims).e.size();
}
int wjb_weight_info(const wjb_dims* dims, int index, char* name_buf, int name_buf_len, size_t* offset, size_t* nbytes, int* dtype) {
    if (!dims_ok(dims)) return set_error("unsupported dims (head dim must be 64)");
    Layout L = build_layout(*dims);
    if (index < 0 || index >= (int)L.e.size()) return set_error("weight index out of range");
    const Entry& e = L.e[index];
    if (name_buf && name_buf_len > 0) {
        strncpy(name_buf, e.name.c_str(), name_buf_len - 1);
        name_buf[name_buf_len - 1] = 0;
    }
    if (offset) *offset = e.offset;
    if (nbytes) *nbytes = e.bytes;
    if (dtype) *dtype = e.dtype;
    return 0;
}

int wjb_model_create(const wjb_dims* dims, const void* weights_blob, wjb_model** out) {
    if (!dims_ok(dims)) return set_error("unsupported dims (need head dim 64, n_mels<=128, ctx<=448/1536)");
    if (!weights_blob || !out) return set_error("null argument");
    if (int e = ensure_init()) return e;
    wjb_model* m = new wjb_model();
    m->d = *dims;
    m->blob = reinterpret_cast<const uint8_t*>(weights_blob);
    m->L = build_layout(*dims);
    if (cudaHostAlloc(&m->h_done, sizeof(int) * 4, cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc(&m->h_ctl, sizeof(DecodeCtl), cudaHostAllocDefault) != cudaSuccess) {
        wjb_model_destroy(m);
        return set_error("cudaHostAlloc failed");
    }
    if (cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev, cudaEventDisableTiming) != cudaSuccess) {
        wjb_model_destroy(m);
        return set_error("stream/event creation failed");
    }
    *out = m;
    return 0;
}

void wjb_model_destroy(wjb_model* m) {
    if (!m) return;
    if (m->graph) cudaGraphExecDestroy(m->graph);
    if (m->h_done) cudaFreeHost(m->h_done);
    if (m->h_ctl) cudaFreeHost(m->h_ctl);
    if (m->own_stream) cudaStreamDestroy(m->own_stream);
    if (m->ev) cudaEventDestroy(m->ev);
    delete m;
}

// ------------------------------------------------------------------ log-mel
size_t wjb_logmel_workspace_bytes(int n_clips, int n_mels) { return (size_t)n_clips * 4 + 256 + (size_t)n_mels * 8 + 256; }

int wjb_logmel_f16(const float* audio, int64_t audio_stride, const int32_t* n_samples, int n_clips, int n_mels, const float* filters,
                   void* out, int time_major, int64_t out_clip_stride, int row0, int n_frames, int reflect_total, void* workspace,
                   void* stream) {
    if (!audio || !n_samples || !filters || !out || !workspace) return set_error("logmel: null argument");
    if (int e = ensure_init()) return e;
    LogmelArgs a;
    a.audio = audio;
    a.audio_stride = audio_stride;
    a.n_samples = n_samples;
    a.n_clips = n_clips;
    a.n_mels = n_mels;
    a.filters = filters;
    a.out = reinterpret_cast<__half*>(out);
    a.time_major = time_major;
    a.out_clip_stride = out_clip_stride;
    a.row0 = row0;
    a.n_frames = n_frames;
    a.reflect_total = reflect_total;
    a.clip_max = reinterpret_cast<float*>(workspace);
    a.mel_range = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(workspace) + (((size_t)n_clips * 4 + 255) & ~size_t(255)));
    return launch_logmel(a, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ encoder
static size_t al(size_t x) { return (x + 255) & ~size_t(255); }

struct EncWs {
    __half *c1, *x, *h, *qkv, *mlp;
    size_t total;
};
static EncWs enc_ws(const wjb_dims& d, int B, uint8_t* base) {
    EncWs w;
    const size_t n = d.n_audio_state, T = d.n_audio_ctx, T2 = 2 * T + 2;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        uint8_t* p = base ? base + off : nullptr;
        off += al(bytes);
        return reinterpret_cast<__half*>(p);
    };
    w.c1 = take((size_t)B * T2 * n * 2);
    w.x = take((size_t)B * T * n * 2);
    w.h = take((size_t)B * T * n * 2);
    w.qkv = take((size_t)B * T * 3 * n * 2);
    w.mlp = take((size_t)B * T * 4 * n * 2);
    w.total = off;
    return w;
}

size_t wjb_encoder_workspace_bytes(const wjb_model* m, int batch) {
    if (!m || batch <= 0) return 0;
    return enc_ws(m->d, batch, nullptr).total;
}

int wjb_encoder_forward(wjb_model* m, const void* mel_tm, int batch, void* out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!m || !mel_tm || !out || !workspace) return set_error("encoder: null argument");
    if (int e = ensure_init()) return e;
    const wjb_dims& d = m->d;
    cudaStream_t s = (cudaStream_t)stream;
    const int B = batch, n = d.n_audio_state, T = d.n_audio_ctx, T2 = 2 * T + 2, C = d.n_mels, H = d.n_audio_head;
    EncWs w = enc_ws(d, B, reinterpret_cast<uint8_t*>(workspace));
    if (w.total > workspace_bytes) return set_error("encoder: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    // zero pad rows 0 and 2T+1 of the conv1 output
    for (int r = 0; r < 2; ++r) {
        cudaError_t e = cudaMemset2DAsync(w.c1 + (size_t)(r ? (T2 - 1) : 0) * n, (size_t)T2 * n * 2, 0, (size_t)n * 2, B, s);
        if (e != cudaSuccess) return set_error("encoder memset: %s", cudaGetErrorString(e));
    }
    GemmArgs g;
    // conv1: [B][2T+2][C] -> rows 1..2T of c1, GELU
    g = GemmArgs();
    g.A = reinterpret_cast<const __half*>(mel_tm);
    g.a_row_stride = C;
    g.a_batch_stride = (long long)T2 * C;
    g.rows_per_batch = 2 * T;
    g.n_batch = B;
    g.K = 3 * C;
    g.W = m->h16("enc.conv1.w");
    g.N = n;
    g.ldw = 3 * C;
    g.bias = m->h16("enc.conv1.b");
    g.out = w.c1 + n;
    g.out_row_stride = n;
    g.out_batch_stride = (long long)T2 * n;
    g.flags = GEMM_GELU;
    {
        ProfScope ps(PC_GEMM, s);
        if (int e = launch_gemm(g, s)) return e;
    }
    // conv2 (stride 2) + GELU + positional embedding -> x [B][T][n]
    g = GemmArgs();
    g.A = w.c1;
    g.a_row_stride = 2 * n;
    g.a_batch_stride = (long long)T2 * n;
    g.rows_per_batch = T;
    g.n_batch = B;
    g.K = 3 * n;
    g.W = m->h16("enc.conv2.w");
    g.N = n;
    g.ldw = 3 * n;
    g.bias = m->h16("enc.conv2.b");
    g.pos = m->f32("enc.pos");
    g.out = w.x;
    g.out_row_stride = n;
    g.out_batch_stride = (long long)T * n;
    g.flags = GEMM_GELU;
    {
        ProfScope ps(PC_GEMM, s);
        if (int e = launch_gemm(g, s)) return e;
    }

    const int M = B * T;
    auto linear = [&](const __half* A, int K, const __half* W, const __half* bias, const __half* res, __half* o, int N, int flags) {
        GemmArgs q;
        q.A = A;
        q.a_row_stride = K;
        q.rows_per_batch = M;
        q.n_batch = 1;
        q.K = K;
        q.W = W;
        q.N = N;
        q.ldw = K;
        q.bias = bias;
        q.residual = res;
        q.out = o;
        q.out_row_stride = N;
        q.flags = flags;
        ProfScope ps(PC_GEMM, s);
        return launch_gemm(q, s);
    };
    auto ln = [&](const __half* x, const __half* g_, const __half* b_, __half* o) {
        ProfScope ps(PC_LN, s);
        return launch_layernorm(x, g_, b_, o, M, n, s);
    };
    for (int i = 0; i < d.n_audio_layer; ++i) {
        const std::string p = "enc." + std::to_string(i) + ".";
        if (int e = ln(w.x, m->h16(p + "ln1.g"), m->h16(p + "ln1.b"), w.h)) return e;
        if (int e = linear(w.h, n, m->h16(p + "qkv.w"), m->h16(p + "qkv.b"), nullptr, w.qkv, 3 * n, 0)) return e;
        {
            ProfScope ps(PC_ATTN, s);
            if (int e = launch_attn_encoder(w.qkv, w.h, B, T, H, s)) return e;
        }
        if (int e = linear(w.h, n, m->h16(p + "out.w"), m->h16(p + "out.b"), w.x, w.x, n, 0)) return e;
        if (int e = ln(w.x, m->h16(p + "ln2.g"), m->h16(p + "ln2.b"), w.h)) return e;
        if (int e = linear(w.h, n, m->h16(p + "fc1.w"), m->h16(p + "fc1.b"), nullptr, w.mlp, 4 * n, GEMM_GELU)) return e;
        if (int e = linear(w.mlp, 4 * n, m->h16(p + "fc2.w"), m->h16(p + "fc2.b"), w.x, w.x, n, 0)) return e;
        if (m->enc_tap && m->enc_tap_every > 0 && (i + 1) % m->enc_tap_every == 0) {
            const size_t bytes = (size_t)M * n * 2;
            cudaError_t ce = cudaMemcpyAsync(reinterpret_cast<uint8_t*>(m->enc_tap) + ((i + 1) / m->enc_tap_every - 1) * bytes, w.x, bytes,
                                             cudaMemcpyDeviceToDevice, s);
            if (ce != cudaSuccess) return set_error("encoder tap: %s", cudaGetErrorString(ce));
        }
    }
    return ln(w.x, m->h16("enc.ln_post.g"), m->h16("enc.ln_post.b"), reinterpret_cast<__half*>(out));
}

int wjb_encoder_set_tap(wjb_model* m, void* out, int every) {
    if (!m) return set_error("encoder_set_tap: null model");
    m->enc_tap = out;
    m->enc_tap_every = out ? every : 0;
    return 0;
}

// ------------------------------------------------------------------ cross K/V
size_t wjb_cross_kv_bytes(const wjb_model* m, int batch) {
    if (!m) return 0;
    return (size_t)m->d.n_text_layer * batch * 2 * m->d.n_text_head * m->d.n_audio_ctx * 64 * 2;
}

int wjb_cross_kv(wjb_model* m, const void* enc_out, int batch, void* kv_out, void* stream) {
    if (!m || !enc_out || !kv_out) return set_error("cross_kv: null argument");
    if (int e = ensure_init()) return e;
    const wjb_dims& d = m->d;
    if (d.n_audio_state != d.n_text_state) return set_error("cross_kv: audio/text widths differ");
    const int n = d.n_text_state, T = d.n_audio_ctx, H = d.n_text_head;
    const size_t per_layer = (size_t)batch * 2 * H * T * 64;
    for (int i = 0; i < d.n_text_layer; ++i) {
        const std::string p = "dec." + std::to_string(i) + ".";
        GemmArgs g;
        g.A = reinterpret_cast<const __half*>(enc_out);
        g.a_row_stride = n;
        g.a_batch_stride = (long long)T * n;
        g.rows_per_batch = T;
        g.n_batch = batch;
        g.K = n;
        g.W = m->h16(p + "ckv.w");
        g.N = 2 * n;
        g.ldw = n;
        g.bias = m->h16(p + "ckv.b");
        g.out = reinterpret_cast<__half*>(kv_out) + (size_t)i * per_layer;
        g.flags = GEMM_HEADSPLIT;
        g.hs_T = T;
        g.hs_H = 2 * H;
        if (int e = launch_gemm(g, (cudaStream_t)stream)) return e;
    }
    return 0;
}

// ------------------------------------------------------------------ greedy decode
struct DecWs {
    __half *x, *h, *qkv, *q, *a, *mlp, *logits, *self_kv;
    DecodeCtl* ctl;
    unsigned char* done;
    long long* lnstat;  // [3 * n_text_layer + 1 LayerNorm sites][B][2] fixed-point row statistics of the residual stream (gemm_step.cu)
    int logits_stride;
    size_t total;
};
static DecWs dec_ws(const wjb_dims& d, int B, uint8_t* base) {
    DecWs w;
    const size_t n = d.n_text_state;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        uint8_t* p = base ? base + off : nullptr;
        off += al(bytes);
        return p;
    };
    w.logits_stride = (d.n_vocab + 63) & ~63;
    w.x = (__half*)take((size_t)B * n * 2);
    w.h = (__half*)take((size_t)B * n * 2);
    w.qkv = (__half*)take((size_t)B * 3 * n * 2);
    w.q = (__half*)take((size_t)B * n * 2);
    w.a = (__half*)take((size_t)B * n * 2);
    w.mlp = (__half*)take((size_t)B * 4 * n * 2);
    w.logits = (__half*)take((size_t)B * w.logits_stride * 2);
    w.self_kv = (__half*)take((size_t)d.n_text_layer * B * 2 * d.n_text_head * d.n_text_ctx * 64 * 2);
    w.ctl = (DecodeCtl*)take(sizeof(DecodeCtl));
    w.done = (unsigned char*)take((size_t)B);
    w.lnstat = (long long*)take((size_t)(3 * d.n_text_layer + 1) * B * 2 * sizeof(long long));
    w.total = off;
    return w;
}

size_t wjb_decode_workspace_bytes(const wjb_model* m, int batch) {
    if (!m || batch <= 0) return 0;
    return dec_ws(m->d, batch, nullptr).total;
}

int wjb_decode_logits_stride(const wjb_model* m) { return m ? ((m->d.n_vocab + 63) & ~63) : 0; }

int wjb_decode_set_trace(wjb_model* m, void* logits_out, size_t logits_out_bytes, int32_t* sampled_out, const int32_t* forced_tokens) {
    if (!m) return set_error("decode_set_trace: null model");
    m->trace_logits = reinterpret_cast<__half*>(logits_out);
    m->trace_logits_bytes = logits_out ? logits_out_bytes : 0;
    m->trace_sampled = sampled_out;
    m->trace_forced = forced_tokens;
    return 0;
}

size_t wjb_align_qk_bytes(const wjb_model* m, int batch, int n_steps) {
    if (!m || batch <= 0 || n_steps <= 0) return 0;
    const wjb_dims& d = m->d;
    return (size_t)batch * (d.n_text_layer - d.n_text_layer / 2) * d.n_text_head * n_steps * d.n_audio_ctx * 2;
}

int wjb_decode_set_align(wjb_model* m, void* qk_out, int n_steps, const int32_t* n_tokens, float* token_prob_out) {
    if (!m) return set_error("decode_set_align: null model");
    m->align_qk = reinterpret_cast<__half*>(qk_out);
    m->align_steps = qk_out ? n_steps : 0;
    m->align_len = qk_out ? n_tokens : nullptr;
    m->align_prob = qk_out ? token_prob_out : nullptr;
    return 0;
}

size_t wjb_align_workspace_bytes(const wjb_model* m, int batch, int n_steps) {
    if (!m || batch <= 0 || n_steps <= 0) return 0;
    const wjb_dims& d = m->d;
    return align_workspace_bytes(batch, (d.n_text_layer - d.n_text_layer / 2) * d.n_text_head, n_steps, d.n_audio_ctx);
}

int wjb_align_dtw(wjb_model* m, const void* qk, int batch, int n_steps, const int32_t* n_tokens, const int32_t* row_begin, const int32_t* n_rows,
                  const int32_t* n_frames2, int medfilt_width, float* matrix, int32_t* jump_frames, void* workspace, size_t workspace_bytes,
                  void* stream) {
    if (!m || !qk || !n_tokens || !row_begin || !n_rows || !n_frames2 || !matrix || !jump_frames || !workspace)
        return set_error("align_dtw: null argument");
    if (int e = ensure_init()) return e;
    const wjb_dims& d = m->d;
    return launch_align(reinterpret_cast<const __half*>(qk), batch, (d.n_text_layer - d.n_text_layer / 2) * d.n_text_head, n_steps, d.n_audio_ctx,
                        n_tokens, row_begin, n_rows, n_frames2, medfilt_width, matrix, jump_frames, workspace, workspace_bytes, (cudaStream_t)stream);
}

// ---- alignment, step 1 as ONE teacher-forced pass over whole sequences (every position a GEMM row)
constexpr int kPrefillLogitRows = 4096;
struct PfWs {
    __half *x, *h, *qkv, *q, *a, *mlp, *logits;
    size_t total;
};
static PfWs pf_ws(const wjb_dims& d, int R, uint8_t* base) {
    PfWs w;
    const size_t n = d.n_text_state;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        uint8_t* p = base ? base + off : nullptr;
        off += al(bytes);
        return reinterpret_cast<__half*>(p);
    };
    w.x = take((size_t)R * n * 2);
    w.h = take((size_t)R * n * 2);
    w.qkv = take((size_t)R * 3 * n * 2);
    w.q = take((size_t)R * n * 2);
    w.a = take((size_t)R * n * 2);
    w.mlp = take((size_t)R * 4 * n * 2);
    w.logits = take((size_t)(R < kPrefillLogitRows ? R : kPrefillLogitRows) * ((d.n_vocab + 63) & ~63) * 2);
    w.total = off;
    return w;
}

size_t wjb_align_prefill_workspace_bytes(const wjb_model* m, int batch, int n_steps) {
    if (!m || batch <= 0 || n_steps <= 0) return 0;
    return pf_ws(m->d, batch * n_steps, nullptr).total;
}

int wjb_align_prefill(wjb_model* m, const void* cross_kv, const int32_t* tokens, int tokens_stride, const int32_t* n_tokens, int batch, int n_steps,
                      int eot, void* qk_out, float* token_prob_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!m || !cross_kv || !tokens || !n_tokens || !qk_out || !workspace) return set_error("align_prefill: null argument");
    if (int e = ensure_init()) return e;
    const wjb_dims& d = m->d;
    const int n = d.n_text_state, H = d.n_text_head, T = d.n_audio_ctx, B = batch, Lp = n_steps, R = B * Lp;
    if (Lp > d.n_text_ctx || Lp > 256) return set_error("align_prefill: %d positions (limit %d)", Lp, d.n_text_ctx < 256 ? d.n_text_ctx : 256);
    if (tokens_stride < Lp) return set_error("align_prefill: tokens_stride < n_steps");
    cudaStream_t s = (cudaStream_t)stream;
    PfWs w = pf_ws(d, R, reinterpret_cast<uint8_t*>(workspace));
    if (w.total > workspace_bytes) return set_error("align_prefill: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    auto linear = [&](const __half* A, int K, const __half* W, const __half* bias, const __half* res, __half* o, int N, int flags, int rows, long long ostride) {
        GemmArgs q;
        q.A = A;
        q.a_row_stride = K;
        q.rows_per_batch = rows;
        q.n_batch = 1;
        q.K = K;
        q.W = W;
        q.N = N;
        q.ldw = K;
        q.bias = bias;
        q.residual = res;
        q.out = o;
        q.out_row_stride = ostride;
        q.flags = flags;
        return launch_gemm(q, s);
    };
    if (int e = launch_prefill_embed(tokens, tokens_stride, n_tokens, m->h16("dec.emb"), m->h16("dec.pos"), w.x, B, Lp, n, s)) return e;
    const size_t cross_per_layer = (size_t)B * 2 * H * T * 64;
    const int sel0 = d.n_text_layer / 2;
    for (int i = 0; i < d.n_text_layer; ++i) {
        const std::string p = "dec." + std::to_string(i) + ".";
        if (int e = launch_layernorm(w.x, m->h16(p + "ln1.g"), m->h16(p + "ln1.b"), w.h, R, n, s)) return e;
        if (int e = linear(w.h, n, m->h16(p + "qkv.w"), m->h16(p + "qkv.b"), nullptr, w.qkv, 3 * n, 0, R, 3 * n)) return e;
        if (int e = launch_prefill_self_attn(w.qkv, w.a, n_tokens, B, Lp, H, s)) return e;
        if (int e = linear(w.a, n, m->h16(p + "out.w"), m->h16(p + "out.b"), w.x, w.x, n, 0, R, n)) return e;
        if (int e = launch_layernorm(w.x, m->h16(p + "ln2.g"), m->h16(p + "ln2.b"), w.h, R, n, s)) return e;
        if (int e = linear(w.h, n, m->h16(p + "cq.w"), m->h16(p + "cq.b"), nullptr, w.q, n, 0, R, n)) return e;
        CrossCapture cap;
        if (i >= sel0) {
            cap.base = reinterpret_cast<__half*>(qk_out) + (long long)(i - sel0) * H * Lp * T;
            cap.b_stride = (long long)(d.n_text_layer - sel0) * H * Lp * T;
            cap.head_stride = (long long)Lp * T;
        }
        if (int e = launch_prefill_cross_attn(w.q, reinterpret_cast<const __half*>(cross_kv) + i * cross_per_layer, w.a, n_tokens, B, Lp, H, T,
                                              cap.base ? &cap : nullptr, s))
            return e;
        if (int e = linear(w.a, n, m->h16(p + "cout.w"), m->h16(p + "cout.b"), w.x, w.x, n, 0, R, n)) return e;
        if (int e = launch_layernorm(w.x, m->h16(p + "ln3.g"), m->h16(p + "ln3.b"), w.h, R, n, s)) return e;
        if (int e = linear(w.h, n, m->h16(p + "fc1.w"), m->h16(p + "fc1.b"), nullptr, w.mlp, 4 * n, GEMM_GELU, R, 4 * n)) return e;
        if (int e = linear(w.mlp, 4 * n, m->h16(p + "fc2.w"), m->h16(p + "fc2.b"), w.x, w.x, n, 0, R, n)) return e;
    }
    if (token_prob_out) {
        if (int e = launch_layernorm(w.x, m->h16("dec.ln.g"), m->h16("dec.ln.b"), w.h, R, n, s)) return e;
        const int ls = (d.n_vocab + 63) & ~63;
        for (int r0 = 0; r0 < R; r0 += kPrefillLogitRows) {
            const int rows = R - r0 < kPrefillLogitRows ? R - r0 : kPrefillLogitRows;
            if (int e = linear(w.h + (size_t)r0 * n, n, m->h16("dec.emb"), nullptr, nullptr, w.logits, d.n_vocab, 0, rows, ls)) return e;
            if (int e = launch_prefill_prob(w.logits, ls, r0, rows, tokens, tokens_stride, n_tokens, token_prob_out, Lp, eot, s)) return e;
        }
    }
    return 0;
}

// Decode-step placement options (wjb_debug_set_decode_flags; measured with scripts/decode_flags_probe.py, profiles/r2_decode_flags.json):
//   1  LayerNorms folded into the consuming step GEMM through the producers' row statistics (gemm_step.cu)   -- no gain: off
//   2  every step GEMM pulls the NEXT Linear's weights into L2 while it runs (cp.async.bulk.prefetch.L2)     -- +2 %: off
//   4  the cross-attention primes its K/V ring before the programmatic-dependent-launch wait                 -- on
//   8  K/V bulk copies carry an L2 evict-first policy (the stream must not push the layer's weights out of L2) -- -1 %: on
static int g_decode_flags = 4 | 8;

// The kernels of one decoder step for B rows on stream s (captured into the step graph).
static int decode_step(wjb_model* m, const DecWs& w, const void* cross_kv, int B, const wjb_decode_opts& o, const uint8_t* suppress_mask,
                       int32_t* tokens, float* slp, float* nsp, int32_t* out_len, cudaStream_t s, const BeamBufs* beam = nullptr) {
    const wjb_dims& d = m->d;
    const int n = d.n_text_state, H = d.n_text_head, T = d.n_audio_ctx;
    // LayerNorm (ln_g / ln_b, may be null) into w.h, then the Linear
    const int chunk = B <= 128 ? B : (B + 1) / 2 <= 128 ? (B + 1) / 2 : 128;
    auto step_ok = [&](int N, int K) { return gemm_step_supported(chunk, N, K) && gemm_step_supported(B - (B - 1) / chunk * chunk, N, K); };
    // site_in: LayerNorm site whose statistics the producers of A accumulated (-1: run the LayerNorm as its own launch);
    // site_out: site this Linear's output rows are the input of (their statistics are added by its epilogue; -1: none)
    auto linear = [&](const __half* A, int K, const __half* W, int ldw, const __half* bias, const __half* res, __half* out, int N,
                      long long out_stride, int flags, const __half* ln_g = nullptr, const __half* ln_b = nullptr, int site_in = -1,
                      int site_out = -1, const __half* next_w = nullptr, size_t next_w_bytes = 0) {
        const bool step_kernel = step_ok(N, K);
        const bool fold_ln = ln_g && site_in >= 0 && step_kernel;
        if (ln_g && !fold_ln) {
            if (int e = launch_layernorm(A, ln_g, ln_b, w.h, B, K, s)) return e;
            A = w.h;
        }
        // the cluster split-K kernel written for exactly this shape class (gemm_step.cu) takes <= 128 rows: larger row counts
        // (64 windows x beam 3 = 192) go through it in row chunks -- the weights stream twice, still well ahead of the general kernel
        if (step_kernel) {
            for (int r0 = 0; r0 < B; r0 += chunk) {
                const int rows = B - r0 < chunk ? B - r0 : chunk;
                StepGemmArgs g;
                g.A = A + (long long)r0 * K;
                g.a_row_stride = K;
                g.rows = rows;
                g.K = K;
                g.W = W;
                g.N = N;
                g.ldw = ldw;
                g.bias = bias;
                g.residual = res ? res + (long long)r0 * out_stride : nullptr;
                g.out = out + (long long)r0 * out_stride;
                g.out_row_stride = out_stride;
                g.flags = flags;
                g.w_constant = true;
                if (fold_ln) {
                    g.ln_gamma = ln_g;
                    g.ln_beta = ln_b;
                    g.ln_stats = w.lnstat + ((long long)site_in * B + r0) * 2;
                }
                if (site_out >= 0) g.out_stats = w.lnstat + ((long long)site_out * B + r0) * 2;
                if ((g_decode_flags & 2) && next_w && r0 == 0) {
                    g.prefetch = next_w;
                    g.prefetch_bytes = next_w_bytes;
                }
                if (int e = launch_gemm_step(g, s)) return e;
            }
            return 0;
        }
        GemmArgs q;
        q.A = A;
        q.a_row_stride = K;
        q.rows_per_batch = B;
        q.n_batch = 1;
        q.K = K;
        q.W = W;
        q.N = N;
        q.ldw = ldw;
        q.bias = bias;
        q.residual = res;
        q.out = out;
        q.out_row_stride = out_stride;
        q.flags = flags;
        q.block_n = 0;  // auto: 32-wide tiles when that still leaves SMs idle, else 64
        return launch_gemm(q, s);
    };
    // beam search: rows = windows x beams; token rows and the cache ancestry are double buffered by step parity, the rows of a
    // window share its cross K/V
    const int kv_div = beam ? beam->beam : 1;
    // the residual stream's row statistics: site 3 i + {0, 1, 2} = input of layer i's attn_ln / cross_attn_ln / mlp_ln.  The embedding
    // kernel writes site 0 and zeroes the others, every Linear that stores the stream adds to the site that reads it next.
    const bool stats_ok = (g_decode_flags & 1) && step_ok(n, n) && step_ok(n, 4 * n);  // every producer of the stream (out, cross-out, fc2) must be a step GEMM
    const int n_sites = stats_ok ? 3 * d.n_text_layer + 1 : 0;
    if (int e = launch_embed(beam ? beam->tokens : tokens, o.tokens_stride, m->h16("dec.emb"), m->h16("dec.pos"), w.x, w.ctl, B, n, s,
                             beam ? beam->tokens_parity_stride : 0, w.lnstat, n_sites))
        return e;
    auto site = [&](int k) { return stats_ok ? k : -1; };
    const size_t self_per_layer = (size_t)B * 2 * H * d.n_text_ctx * 64;
    const size_t cross_per_layer = (size_t)(B / kv_div) * 2 * H * T * 64;
    for (int i = 0; i < d.n_text_layer; ++i) {
        const std::string p = "dec." + std::to_string(i) + ".";
        const size_t nn = (size_t)n * n * 2;  // bytes of an n x n fp16 weight
        const std::string pn = "dec." + std::to_string(i + 1 < d.n_text_layer ? i + 1 : 0) + ".";  // the next step starts at layer 0 again
        if (int e = linear(w.x, n, m->h16(p + "qkv.w"), n, m->h16(p + "qkv.b"), nullptr, w.qkv, 3 * n, 3 * n, 0, m->h16(p + "ln1.g"), m->h16(p + "ln1.b"), site(3 * i),
                           -1, m->h16(p + "out.w"), nn))
            return e;
        if (int e = launch_attn_dec_self(w.qkv, w.self_kv + i * self_per_layer, w.a, &w.ctl->step, w.done, B, H, d.n_text_ctx, s,
                                         beam ? beam->anc : nullptr, beam ? beam->anc_parity_stride : 0)) return e;
        if (int e = linear(w.a, n, m->h16(p + "out.w"), n, m->h16(p + "out.b"), w.x, w.x, n, n, 0, nullptr, nullptr, -1, site(3 * i + 1), m->h16(p + "cq.w"), nn))
            return e;
        if (int e = linear(w.x, n, m->h16(p + "cq.w"), n, m->h16(p + "cq.b"), nullptr, w.q, n, n, 0, m->h16(p + "ln2.g"), m->h16(p + "ln2.b"), site(3 * i + 1))) return e;
        CrossCapture cap;
        const int sel0 = d.n_text_layer / 2;  // alignment heads: every head of the upper half of the layers (upstream default)
        if (m->align_qk && !beam && i >= sel0) {
            cap.base = m->align_qk + (long long)(i - sel0) * H * m->align_steps * T;
            cap.b_stride = (long long)(d.n_text_layer - sel0) * H * m->align_steps * T;
            cap.head_stride = (long long)m->align_steps * T;
            cap.step = &w.ctl->step;
        }
        CrossTuning tune;
        tune.early_kv = (g_decode_flags & 4) ? 1 : 0;
        tune.evict_first = (g_decode_flags & 8) ? 1 : 0;
        if (g_decode_flags & 2) {
            tune.pf_ptr = m->h16(p + "cout.w");
            tune.pf_bytes = nn;
        }
        if (int e = launch_attn_dec_cross(w.q, reinterpret_cast<const __half*>(cross_kv) + i * cross_per_layer, w.a, w.done, B, H, T, s, kv_div,
                                          cap.base ? &cap : nullptr, &tune))
            return e;
        if (int e = linear(w.a, n, m->h16(p + "cout.w"), n, m->h16(p + "cout.b"), w.x, w.x, n, n, 0, nullptr, nullptr, -1, site(3 * i + 2), m->h16(p + "fc1.w"), 4 * nn))
            return e;
        if (int e = linear(w.x, n, m->h16(p + "fc1.w"), n, m->h16(p + "fc1.b"), nullptr, w.mlp, 4 * n, 4 * n, GEMM_GELU, m->h16(p + "ln3.g"), m->h16(p + "ln3.b"), site(3 * i + 2),
                           -1, m->h16(p + "fc2.w"), 4 * nn))
            return e;
        // the last layer's output goes through ln (dec.ln) in front of the logits GEMM, which is not a step GEMM: no statistics needed
        if (int e = linear(w.mlp, 4 * n, m->h16(p + "fc2.w"), 4 * n, m->h16(p + "fc2.b"), w.x, w.x, n, n, 0, nullptr, nullptr, -1,
                           i + 1 < d.n_text_layer ? site(3 * i + 3) : -1, m->h16(pn + "qkv.w"), 3 * nn))
            return e;
    }
    if (int e = linear(w.x, n, m->h16("dec.emb"), n, nullptr, nullptr, w.logits, d.n_vocab, w.logits_stride, 0, m->h16("dec.ln.g"), m->h16("dec.ln.b"))) return e;
    DecodeParams p;
    p.B = B;
    p.n_vocab = d.n_vocab;
    p.logits_stride = w.logits_stride;
    p.eot = o.eot;
    p.no_speech = o.no_speech;
    p.no_timestamps = o.no_timestamps;
    p.timestamp_begin = o.timestamp_begin;
    p.suppress_blank = o.suppress_blank;
    p.blank_token = o.blank_token;
    p.apply_timestamp_rules = o.apply_timestamp_rules;
    p.max_initial_timestamp_index = o.max_initial_timestamp_index;
    p.n_ctx = d.n_text_ctx;
    p.tokens_stride = o.tokens_stride;
    if (beam) {
        if (int e = launch_beam_select(w.logits, suppress_mask, *beam, nsp, w.done, w.ctl, p, s)) return e;
    } else {
        p.trace_logits = m->trace_logits;
        p.trace_sampled = m->trace_sampled;
        p.trace_forced = m->trace_forced;
        p.align_len = m->align_len;
        p.align_prob = m->align_prob;
        if (int e = launch_sample(w.logits, suppress_mask, tokens, slp, nsp, out_len, w.done, w.ctl, p, s)) return e;
    }
    return launch_advance(w.ctl, s);
}

int wjb_decode_greedy(wjb_model* m, const void* cross_kv, int batch, const wjb_decode_opts* opts, const uint8_t* suppress_mask,
                      int32_t* tokens, float* sum_logprob, float* no_speech_prob, int32_t* out_len, void* workspace,
                      size_t workspace_bytes, int* steps_run, void* stream) {
    if (!m || !cross_kv || !opts || !tokens || !sum_logprob || !no_speech_prob || !out_len || !workspace)
        return set_error("decode: null argument");
    if (int e = ensure_init()) return e;
    const wjb_dims& d = m->d;
    // order after the caller's stream, then run on our own capturable stream; synchronous on return
    cudaStream_t s = m->own_stream;
    cudaEventRecord(m->ev, (cudaStream_t)stream);
    cudaStreamWaitEvent(s, m->ev, 0);
    const wjb_decode_opts& o = *opts;
    wjb_decode_opts okey = o;  // graph identity: everything except the run-time sampling knobs
    okey.temperature = 0.f;
    okey.seed = 0;
    if (o.n_initial < 1 || o.sample_len < 1) return set_error("decode: bad n_initial/sample_len");
    const int total_steps = o.n_initial - 1 + o.sample_len;
    if (total_steps > d.n_text_ctx) return set_error("decode: n_initial + sample_len exceeds n_text_ctx");
    if (o.tokens_stride < o.n_initial + o.sample_len) return set_error("decode: tokens_stride too small");
    DecWs w = dec_ws(d, batch, reinterpret_cast<uint8_t*>(workspace));
    if (w.total > workspace_bytes) return set_error("decode: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    if (m->align_qk && (!m->trace_forced || !m->align_len)) return set_error("decode: the alignment pass needs forced tokens and their lengths");
    if (m->align_qk && total_steps > m->align_steps) return set_error("decode: alignment capture holds %d steps, the run has %d", m->align_steps, total_steps);
    if (m->trace_logits && m->trace_logits_bytes < (size_t)total_steps * batch * w.logits_stride * 2)
        return set_error("decode: logits trace buffer too small (%zu < %zu)", m->trace_logits_bytes, (size_t)total_steps * batch * w.logits_stride * 2);

    // reset per-run state
    DecodeCtl* h_ctl = m->h_ctl;  // pinned: the async copy below may outlive this frame's locals
    h_ctl->step = 0;
    h_ctl->n_initial = o.n_initial;
    h_ctl->sot_index = o.sot_index;
    h_ctl->max_steps = o.sample_len;
    h_ctl->n_done = 0;
    h_ctl->temperature = o.temperature;
    h_ctl->seed = o.seed;
    cudaError_t ce;
    if ((ce = cudaMemcpyAsync(w.ctl, h_ctl, sizeof(*h_ctl), cudaMemcpyHostToDevice, s)) != cudaSuccess)
        return set_error("decode ctl copy: %s", cudaGetErrorString(ce));
    cudaMemsetAsync(w.done, 0, batch, s);
    cudaMemsetAsync(sum_logprob, 0, sizeof(float) * batch, s);
    cudaMemsetAsync(no_speech_prob, 0, sizeof(float) * batch, s);
    cudaMemsetAsync(out_len, 0, sizeof(int32_t) * batch, s);
    const bool hit = m->graph && m->g_kv == cross_kv && m->g_ws == workspace && m->g_B == batch && m->g_mask == suppress_mask &&
                     m->g_tokens == tokens && m->g_slp == sum_logprob && m->g_nsp == no_speech_prob && m->g_len == out_len &&
                     m->g_trace_logits == m->trace_logits && m->g_trace_sampled == m->trace_sampled && m->g_trace_forced == m->trace_forced &&
                     m->g_align_qk == m->align_qk && m->g_align_len == m->align_len && m->g_align_prob == m->align_prob &&
                     m->g_align_steps == m->align_steps && m->g_flags == g_decode_flags &&
                     memcmp(&m->g_opts, &okey, sizeof(okey)) == 0;
    if (!hit) {
        if (m->graph) {
            cudaGraphExecDestroy(m->graph);
            m->graph = nullptr;
        }
        cudaStreamSynchronize(s);
        cudaGraph_t graph = nullptr;
        if ((ce = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal)) != cudaSuccess)
            return set_error("decode: begin capture: %s", cudaGetErrorString(ce));
        set_pdl(true);
        int e = decode_step(m, w, cross_kv, batch, o, suppress_mask, tokens, sum_logprob, no_speech_prob, out_len, s);
        set_pdl(false);
        ce = cudaStreamEndCapture(s, &graph);
        if (e) {
            if (graph) cudaGraphDestroy(graph);
            return e;
        }
        if (ce != cudaSuccess) return set_error("decode: end capture: %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&m->graph, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) return set_error("decode: graph instantiate: %s", cudaGetErrorString(ce));
        m->g_kv = cross_kv;
        m->g_ws = workspace;
        m->g_B = batch;
        m->g_mask = suppress_mask;
        m->g_tokens = tokens;
        m->g_slp = sum_logprob;
        m->g_nsp = no_speech_prob;
        m->g_len = out_len;
        m->g_opts = okey;
        m->g_trace_logits = m->trace_logits;
        m->g_trace_sampled = m->trace_sampled;
        m->g_trace_forced = m->trace_forced;
        m->g_align_qk = m->align_qk;
        m->g_align_len = m->align_len;
        m->g_align_prob = m->align_prob;
        m->g_align_steps = m->align_steps;
        m->g_flags = g_decode_flags;
    }
    const int check_every = o.check_every > 0 ? o.check_every : 8;
    int step = 0;
    for (; step < total_steps; ++step) {
        if ((ce = cudaGraphLaunch(m->graph, s)) != cudaSuccess) return set_error("decode: graph launch: %s", cudaGetErrorString(ce));
        if ((step + 1) % check_every == 0 && step + 1 >= o.n_initial) {
            cudaMemcpyAsync(m->h_done, &w.ctl->n_done, sizeof(int), cudaMemcpyDeviceToHost, s);
            if ((ce = cudaStreamSynchronize(s)) != cudaSuccess) return set_error("decode: sync: %s", cudaGetErrorString(ce));
            if (*m->h_done >= batch) {
                ++step;
                break;
            }
        }
    }
    if ((ce = cudaStreamSynchronize(s)) != cudaSuccess) return set_error("decode: final sync: %s", cudaGetErrorString(ce));
    if (steps_run) *steps_run = step;
    return 0;
}

// ------------------------------------------------------------------ profiling
void wjb_profile_enable(int on) { g_prof_on = on != 0; }

int wjb_profile_read(float* ms_by_class, int* launches_by_class, int n_classes) {
    for (int i = 0; i < n_classes; ++i) {
        if (ms_by_class) ms_by_class[i] = 0.f;
        if (launches_by_class) launches_by_class[i] = 0;
    }
    for (auto& r : g_prof) {
        cudaEventSynchronize(r.b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess && r.cls < n_classes) {
            if (ms_by_class) ms_by_class[r.cls] += ms;
            if (launches_by_class) launches_by_class[r.cls] += 1;
        }
        g_ev_pool.push_back(r.a);
        g_ev_pool.push_back(r.b);
    }
    g_prof.clear();
    return 0;
}

// ------------------------------------------------------------------ beam search decode
int wjb_decode_beam(wjb_model* m, const void* cross_kv, const wjb_beam_bufs* bufs, const wjb_decode_opts* opts, const uint8_t* suppress_mask,
                    float* no_speech_prob, void* workspace, size_t workspace_bytes, int* steps_run, void* stream) {
    if (!m || !cross_kv || !bufs || !opts || !no_speech_prob || !workspace) return set_error("decode_beam: null argument");
    if (!bufs->tokens || !bufs->anc || !bufs->sum_logprob || !bufs->fin_tokens || !bufs->fin_score || !bufs->fin_len || !bufs->fin_count ||
        !bufs->audio_done)
        return set_error("decode_beam: null buffer");
    if (int e = ensure_init()) return e;
    const wjb_dims& d = m->d;
    const wjb_decode_opts& o = *opts;
    const int n_audio = bufs->n_audio, beam = bufs->beam_size, rows = n_audio * beam;
    if (n_audio < 1 || beam < 1 || beam > kMaxBeam) return set_error("decode_beam: beam_size %d outside 1..%d", beam, kMaxBeam);
    if (bufs->max_candidates < 1 || bufs->max_candidates > kMaxBeam * 4) return set_error("decode_beam: max_candidates %d", bufs->max_candidates);
    if (o.temperature != 0.f) return set_error("decode_beam: beam search runs at temperature 0");
    if (o.n_initial < 1 || o.sample_len < 1) return set_error("decode_beam: bad n_initial/sample_len");
    const int total_steps = o.n_initial - 1 + o.sample_len;
    if (total_steps > d.n_text_ctx) return set_error("decode_beam: n_initial + sample_len exceeds n_text_ctx");
    if (o.tokens_stride < o.n_initial + o.sample_len) return set_error("decode_beam: tokens_stride too small");
    cudaStream_t s = m->own_stream;
    cudaEventRecord(m->ev, (cudaStream_t)stream);
    cudaStreamWaitEvent(s, m->ev, 0);
    DecWs w = dec_ws(d, rows, reinterpret_cast<uint8_t*>(workspace));
    if (w.total > workspace_bytes) return set_error("decode_beam: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    BeamBufs bb;
    bb.n_audio = n_audio;
    bb.beam = beam;
    bb.rows = rows;
    bb.max_candidates = bufs->max_candidates;
    bb.tokens = bufs->tokens;
    bb.tokens_parity_stride = (long long)rows * o.tokens_stride;
    bb.anc = bufs->anc;
    bb.anc_parity_stride = (long long)rows * d.n_text_ctx;
    bb.sum_logprob = bufs->sum_logprob;
    bb.fin_tokens = bufs->fin_tokens;
    bb.fin_score = bufs->fin_score;
    bb.fin_len = bufs->fin_len;
    bb.fin_count = bufs->fin_count;
    bb.audio_done = bufs->audio_done;

    DecodeCtl* h_ctl = m->h_ctl;
    h_ctl->step = 0;
    h_ctl->n_initial = o.n_initial;
    h_ctl->sot_index = o.sot_index;
    h_ctl->max_steps = o.sample_len;
    h_ctl->n_done = 0;
    h_ctl->temperature = 0.f;
    h_ctl->seed = 0;
    cudaError_t ce;
    if ((ce = cudaMemcpyAsync(w.ctl, h_ctl, sizeof(*h_ctl), cudaMemcpyHostToDevice, s)) != cudaSuccess)
        return set_error("decode_beam ctl copy: %s", cudaGetErrorString(ce));
    cudaMemsetAsync(w.done, 0, rows, s);
    cudaMemsetAsync(no_speech_prob, 0, sizeof(float) * n_audio, s);
    if ((ce = cudaStreamSynchronize(s)) != cudaSuccess) return set_error("decode_beam: setup sync: %s", cudaGetErrorString(ce));

    // one step as a CUDA graph (re-captured per call: its identity would be a dozen pointers and a run is hundreds of replays)
    cudaGraphExec_t exec = nullptr;
    {
        cudaGraph_t graph = nullptr;
        if ((ce = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal)) != cudaSuccess)
            return set_error("decode_beam: begin capture: %s", cudaGetErrorString(ce));
        set_pdl(true);
        int e = decode_step(m, w, cross_kv, rows, o, suppress_mask, nullptr, nullptr, no_speech_prob, nullptr, s, &bb);
        set_pdl(false);
        ce = cudaStreamEndCapture(s, &graph);
        if (e) {
            if (graph) cudaGraphDestroy(graph);
            return e;
        }
        if (ce != cudaSuccess) return set_error("decode_beam: end capture: %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) return set_error("decode_beam: graph instantiate: %s", cudaGetErrorString(ce));
    }
    const int check_every = o.check_every > 0 ? o.check_every : 8;
    int step = 0, rc = 0;
    for (; step < total_steps; ++step) {
        if ((ce = cudaGraphLaunch(exec, s)) != cudaSuccess) {
            rc = set_error("decode_beam: graph launch: %s", cudaGetErrorString(ce));
            break;
        }
        if ((step + 1) % check_every == 0 && step + 1 >= o.n_initial) {
            cudaMemcpyAsync(m->h_done, &w.ctl->n_done, sizeof(int), cudaMemcpyDeviceToHost, s);
            if ((ce = cudaStreamSynchronize(s)) != cudaSuccess) {
                rc = set_error("decode_beam: sync: %s", cudaGetErrorString(ce));
                break;
            }
            if (*m->h_done >= rows) {
                ++step;
                break;
            }
        }
    }
    if (!rc && (ce = cudaStreamSynchronize(s)) != cudaSuccess) rc = set_error("decode_beam: final sync: %s", cudaGetErrorString(ce));
    if (exec) cudaGraphExecDestroy(exec);
    if (steps_run) *steps_run = step;
    return rc;
}

// ------------------------------------------------------------------ building blocks
int wjb_gemm_f16(const void* A, int64_t a_row_stride, int64_t a_batch_stride, int rows_per_batch, int n_batch, int K, const void* W,
                 int N, int ldw, const void* bias, const void* residual, void* out, int64_t out_row_stride, int64_t out_batch_stride,
                 int flags, int block_n, void* stream) {
    if (int e = ensure_init()) return e;
    GemmArgs g;
    g.A = (const __half*)A;
    g.a_row_stride = a_row_stride;
    g.a_batch_stride = a_batch_stride;
    g.rows_per_batch = rows_per_batch;
    g.n_batch = n_batch;
    g.K = K;
    g.W = (const __half*)W;
    g.N = N;
    g.ldw = ldw;
    g.bias = (const __half*)bias;
    g.residual = (const __half*)residual;
    g.out = (__half*)out;
    g.out_row_stride = out_row_stride;
    g.out_batch_stride = out_batch_stride;
    g.flags = flags & GEMM_GELU;
    g.block_n = block_n;
    return launch_gemm(g, (cudaStream_t)stream);
}

constexpr size_t kSplitKBytes = 8u << 20;
size_t wjb_gemm_splitk_workspace_bytes(void) { return kSplitKCounters * sizeof(unsigned) + kSplitKBytes; }

void wjb_debug_gemm_trace(void* buf) {
    gemm_set_trace(buf);
    gemm_step_set_trace(buf);
}

int wjb_gemm_step_ln_f16(const void* A, int64_t a_row_stride, int rows, int K, const void* ln_gamma, const void* ln_beta, const void* W, int N,
                         int ldw, const void* bias, const void* residual, void* out, int64_t out_row_stride, int flags, int block_n,
                         int cluster, int w_constant, void* stream) {
    return wjb_gemm_step_stats_f16(A, a_row_stride, rows, K, ln_gamma, ln_beta, nullptr, W, N, ldw, bias, residual, out, out_row_stride, nullptr, flags,
                                   block_n, cluster, w_constant, stream);
}

int wjb_gemm_step_stats_f16(const void* A, int64_t a_row_stride, int rows, int K, const void* ln_gamma, const void* ln_beta, const int64_t* ln_stats,
                            const void* W, int N, int ldw, const void* bias, const void* residual, void* out, int64_t out_row_stride,
                            int64_t* out_stats, int flags, int block_n, int cluster, int w_constant, void* stream) {
    if (int e = ensure_init()) return e;
    StepGemmArgs g;
    g.ln_stats = reinterpret_cast<const long long*>(ln_stats);
    g.out_stats = reinterpret_cast<long long*>(out_stats);
    g.A = (const __half*)A;
    g.a_row_stride = a_row_stride;
    g.rows = rows;
    g.K = K;
    g.W = (const __half*)W;
    g.N = N;
    g.ldw = ldw;
    g.bias = (const __half*)bias;
    g.residual = (const __half*)residual;
    g.out = (__half*)out;
    g.out_row_stride = out_row_stride;
    g.flags = flags & GEMM_GELU;
    g.block_n = block_n;
    g.cluster = cluster;
    g.w_constant = w_constant != 0;
    g.ln_gamma = (const __half*)ln_gamma;
    g.ln_beta = (const __half*)ln_beta;
    return launch_gemm_step(g, (cudaStream_t)stream);
}

int wjb_gemm_step_f16(const void* A, int64_t a_row_stride, int rows, int K, const void* W, int N, int ldw, const void* bias,
                      const void* residual, void* out, int64_t out_row_stride, int flags, int block_n, int cluster, int w_constant,
                      void* stream) {
    return wjb_gemm_step_ln_f16(A, a_row_stride, rows, K, nullptr, nullptr, W, N, ldw, bias, residual, out, out_row_stride, flags, block_n, cluster,
                                w_constant, stream);
}
void wjb_debug_set_pdl(int on) { set_pdl(on != 0); }
void wjb_debug_set_decode_flags(int flags) { g_decode_flags = flags; }
int wjb_debug_get_decode_flags(void) { return g_decode_flags; }

int wjb_gemm_f16_splitk(const void* A, int64_t a_row_stride, int rows, int K, const void* W, int N, int ldw, const void* bias,
                        const void* residual, void* out, int64_t out_row_stride, int flags, int block_n, int splits, void* workspace,
                        size_t workspace_bytes, void* stream) {
    if (int e = ensure_init()) return e;
    if (!workspace || workspace_bytes < kSplitKCounters * sizeof(unsigned) + (1u << 20)) return set_error("gemm split-K: workspace too small");
    GemmArgs g;
    g.A = (const __half*)A;
    g.a_row_stride = a_row_stride;
    g.rows_per_batch = rows;
    g.n_batch = 1;
    g.K = K;
    g.W = (const __half*)W;
    g.N = N;
    g.ldw = ldw;
    g.bias = (const __half*)bias;
    g.residual = (const __half*)residual;
    g.out = (__half*)out;
    g.out_row_stride = out_row_stride;
    g.flags = flags & GEMM_GELU;
    g.block_n = block_n;
    g.splits = splits;
    g.splitk_cnt = reinterpret_cast<unsigned*>(workspace);
    g.splitk_ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + kSplitKCounters * sizeof(unsigned));
    g.splitk_ws_bytes = workspace_bytes - kSplitKCounters * sizeof(unsigned);
    return launch_gemm(g, (cudaStream_t)stream);
}

int wjb_frame_head_f16(const void* x, const void* w, float bias, float* prob, int rows, int n, void* stream) {
    if (!x || !w || !prob) return set_error("frame_head: null argument");
    return launch_frame_head((const __half*)x, (const __half*)w, bias, prob, rows, n, (cudaStream_t)stream);
}

int wjb_layernorm_f16(const void* x, const void* gamma, const void* beta, void* out, int rows, int n, void* stream) {
    if (int e = ensure_init()) return e;
    return launch_layernorm((const __half*)x, (const __half*)gamma, (const __half*)beta, (__half*)out, rows, n, (cudaStream_t)stream);
}

int wjb_attention_encoder_f16(const void* qkv, void* out, int batch, int T, int n_head, void* stream) {
    if (int e = ensure_init()) return e;
    return launch_attn_encoder((const __half*)qkv, (__half*)out, batch, T, n_head, (cudaStream_t)stream);
}

int wjb_attention_self_f16(const void* qkv, void* kv_cache, void* out, const int32_t* position, int batch, int n_head, int n_ctx, void* stream) {
    if (int e = ensure_init()) return e;
    return launch_attn_dec_self((const __half*)qkv, (__half*)kv_cache, (__half*)out, position, nullptr, batch, n_head, n_ctx, (cudaStream_t)stream);
}

int wjb_attention_cross_f16(const void* q, const void* kv, void* out, int batch, int n_head, int T, void* stream) {
    if (int e = ensure_init()) return e;
    return launch_attn_dec_cross((const __half*)q, (const __half*)kv, (__half*)out, nullptr, batch, n_head, T, (cudaStream_t)stream);
}

int wjb_attention_cross_beam_f16(const void* q, const void* kv, void* out, int rows, int n_head, int T, int beams, void* stream) {
    if (int e = ensure_init()) return e;
    if (beams < 1 || rows % beams) return set_error("attention_cross_beam: %d rows is not a multiple of %d beams", rows, beams);
    return launch_attn_dec_cross((const __half*)q, (const __half*)kv, (__half*)out, nullptr, rows, n_head, T, (cudaStream_t)stream, beams);
}

}  // extern "C"
