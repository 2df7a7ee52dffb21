// libvapb200: context, weight loading, step orchestration and the C ABI
// declared in include/vapb200.h.
#include "../../include/vapb200.h"
#include "common.cuh"
#include "gemm_tc.cuh"
#include "fused_tf.cuh"
#include "fused_tf2.cuh"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace vapb;

namespace {

thread_local std::string g_create_error;

struct HostTensor {
    const float* data;
    int ndim;
    uint32_t dims[4];
    size_t numel;
};

struct ConvLayer {           // conv1..conv4 as GEMMs over channels-last activations
    int k, s, p, Lin, Lout;
    float* W;                // [256][k*256], K index = tap*256 + cin
    float* b;
    float* cnw;
    float* cnb;
    TcWeight tc;             // bf16 hi/lo planes of W for the tcgen05 path
};

struct AttnWeights {
    float* Wqkv;             // self: [768][256] = [Wq; Wk; Wv]
    float* Wproj;            // [256][256]
    float* slopes;           // [4]
    TcWeight tc_qkv, tc_proj;
    TcWeight tc_q, tc_kv;    // rows [0,256) and [256,768) of Wqkv, used when the layer is pruned to the last frame
};
struct LayerWeights {
    float *ln_sa_w, *ln_sa_b, *ln_ff_w, *ln_ff_b;
    AttnWeights sa;
    float* W1;               // [768][256]
    float* W2;               // [256][768]
    TcWeight tc_w1, tc_w2;
    // cross attention (ar.layers.*) only
    bool cross;
    float *ln_src_w, *ln_src_b;
    float* Wq_c;             // [256][256]
    float* Wkv_c;            // [512][256] = [Wk; Wv]
    float* Wproj_c;
    float* slopes_c;
    TcWeight tc_q_c, tc_kv_c, tc_proj_c;
};

struct GraphEntry {          // one captured step per batch size (audio / out are read through IoPtrs in device memory)
    int B;
    cudaGraphExec_t exec;
    int launches;
    unsigned long long last_used;
};
constexpr size_t kMaxGraphs = 96;      // LRU beyond that (a server's batch size wanders between 1 and max_batch)

}  // namespace

struct vapb_ctx {
    int device = 0;
    int frame_hz = 20, T = 50, max_streams = 0, max_batch = 0, head_kind = 0;
    int S = 1120;                    // samples per chunk
    int L[5] = {0, 0, 0, 0, 0};      // conv output lengths
    int n_lstm = 5;                  // frames fed to the LSTM (= downsample taps)
    std::string err;

    std::vector<void*> allocs;       // everything cudaMalloc'ed, freed in destroy

    // weights
    float *w0 = nullptr, *b0 = nullptr, *cn0w = nullptr, *cn0b = nullptr;
    ConvLayer conv[4];
    float *Wih = nullptr, *Whh = nullptr, *b_lstm = nullptr;
    TcWeight tc_ih;                  // W_ih planes: the LSTM input projection on tcgen05 (option lstm_x_tc); the recurrence stays fp32
    int opt_lstm_x_tc = 1;           // measured: -11 us per step at B = 64, golden file 1.73e-5 (fp32 projection: 1.97e-5)
    float *Wds = nullptr, *bds = nullptr, *ds_lnw = nullptr, *ds_lnb = nullptr;
    TcWeight tc_ds;
    LayerWeights layers[4];          // [0] = ar_channel.layers.0, [1..3] = ar.layers.0..2
    float *Wa = nullptr, *Wb = nullptr, *comb_lnw = nullptr, *comb_lnb = nullptr;
    float *Wh = nullptr, *bh = nullptr;
    int n_out = 256;
    // k-major transposes for the fused newest-frame tail (k_tail): last cross layer + combinator + head
    float *t_WqT = nullptr, *t_WprojT = nullptr, *t_WqcT = nullptr, *t_WprojcT = nullptr, *t_W1T = nullptr, *t_W2T = nullptr,
          *t_WaT = nullptr, *t_WbT = nullptr, *t_WhT = nullptr;
    int opt_tail = 1;                // 1 = k_tail (one kernel), 0 = the nine per-op kernels
    int opt_conv12_ks = 0;
    // L2 warm-up of the weights behind the encoder (side branch of the step graph)
    const void** pf_ptrs = nullptr;
    unsigned long long* pf_bytes = nullptr;
    int pf_n = 0;
    unsigned long long pf_lines = 0;
    int opt_prefetch = 0;            // measured: 0.586 vs 0.579 ms per step in the flushed bench: no gain, off           // experiment: split-K of conv1 (low nibble) and conv2 (high nibble); 0 = none
    float *va_w = nullptr, *va_b = nullptr;

    // per-stream state
    float *hS = nullptr, *cS = nullptr, *ring = nullptr;
    int* count = nullptr;
    int* bulk_ids = nullptr;         // device int = max_streams: the scratch LSTM slot of vapb_score_offline
    int* bulk_count = nullptr;       // scratch frame counters for its window batches

    // layer-0 Q/K/V cache (batched path): LN(e_j) Wq / Wk / Wv of ar_channel depends on frame j only (no positional input,
    // ALiBi is a shift-invariant key bias: modules.py:93-95, 170-212), so it is projected once when the frame arrives and
    // kept in a per-stream ring next to the embedding ring; host-side frame counters know which streams are up to date
    float* qkv_ring = nullptr;       // [max_streams][2][T][768]
    float* QKVn = nullptr;           // [2 max_batch][768] projections of the newest frame
    std::vector<long long> h_cnt, h_qkv;      // frames seen per stream / frames covered by its cached rows (-1 = stale)
    int opt_qkv_cache = 1;
    // per-step workspaces (sized for max_batch)
    uint8_t* iobuf_dev = nullptr;    // [IoPtrs (16 B)][ids: max_batch x int32], refreshed by ONE H2D copy per step
    uint8_t* iobuf_pinned = nullptr;
    IoPtrs* io_dev = nullptr;
    int* ids_dev = nullptr;
    int sm_count = 148;
    unsigned long long tick = 0;
    int* tvalid = nullptr;
    float* act[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // channels-last, with halo rows
    int halo[5] = {0, 0, 0, 0, 0};
    float *hW = nullptr, *cW = nullptr, *Gx = nullptr, *Gt = nullptr, *Y = nullptr, *dsout = nullptr, *ebuf = nullptr;
    float *X = nullptr, *Z = nullptr, *QKV = nullptr, *O = nullptr, *Hd = nullptr, *KVc = nullptr, *Qc = nullptr;
    float* part = nullptr;           // split-K partial sums [kMaxSplit][rows][256]
    size_t part_stride = 0;
    float *Xl = nullptr, *Zl = nullptr, *Ql = nullptr, *Ol = nullptr, *Hl = nullptr;   // one row per sequence (pruned layer)
    float* audio_stage = nullptr;    // device staging for vapb_step_host
    float* out_stage = nullptr;
    TcWorkspace tcws;                // bf16 hi/lo activation planes for the tcgen05 path
    // per-stream persistent transformer kernel (fused_tf.cu): op list in device memory
    FOp* fops = nullptr;
    int n_fops = 0;
    long long* fused_clk = nullptr;  // optional clock64 stamps per op (option fused_dbg)
    int opt_fused = 1, opt_fused_dbg = 0;
    // stream kernel v2 (fused_tf2.cu, T <= 64): LayerNorm-folded concatenated weights, activation planes, op list
    struct V2Layer {
        float *Wg1 = nullptr, *Wqc = nullptr, *W1s = nullptr;          // fp32 [N][256] device copies (sources of the planes)
        float *s_g1 = nullptr, *c_g1 = nullptr, *s_qc = nullptr, *c_qc = nullptr, *s_1 = nullptr, *c_1 = nullptr;
        int n_g1 = 0, nln_g1 = 0;
        TcWeight g1, qc, w1;
    } v2[4];
    float *X2f = nullptr, *St2 = nullptr;
    __nv_bfloat16 *X2h = nullptr, *X2l = nullptr, *G1h = nullptr, *G1l = nullptr;
    F2Op* f2ops = nullptr;
    int n_f2ops = 0;
    int opt_fused_v = 2;             // 2 = second-generation stream kernel where it applies (T <= 64; measured 581 vs 597 us per step at B = 64), 1 = first generation
    const float* fused_ds_part = nullptr;      // downsample partials handed to the stream kernel (its gather op finishes the embedding)
    long long fused_ds_stride = 0;
    int fused_ds_nsplit = 0;

    // taps
    std::map<std::string, std::pair<float*, size_t>> taps;
    float *tap_chan = nullptr, *tap_cross[3] = {nullptr, nullptr, nullptr}, *tap_comb = nullptr, *tap_logits = nullptr,
          *tap_xin = nullptr;
    int last_B = 0;

    // options
    int opt_graph = 1, opt_gemm = 0, opt_keep_taps = 0, opt_timing = 0, opt_lstm_fused = 1, opt_tile_n = 0, opt_fuse_ln = 1, opt_prune = 1, opt_attn_rk = 0, opt_fork = 0, opt_splitk = 1, opt_conv4p = 1, opt_pdl = 0;   // conv4p: 1 = two accumulators, 3 = + lo*lo product   // attn_rk / fork / pdl measured no better (profiles/r01_h_option_ablation.log)   // PDL measured slower inside CUDA graphs on this driver (ablation in profiles/)
    std::vector<GraphEntry> graphs;
    int launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    // CUDA graphs cannot be captured on the legacy default stream: work submitted on it is
    // bridged onto this internal stream with events (stream order is preserved for the caller).
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    // side branch of the step graph: the cross-attention K/V projection only needs the layer input,
    // so it runs next to the LN -> QKV -> attention -> proj chain instead of in front of it
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

namespace {

int fail(vapb_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CK(c, call)                                                                                    \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail((c), VAPB_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),       \
                        __FILE__, __LINE__);                                                           \
    } while (0)

// ---- VAPW blob parsing (vap_realtime_b200/weights.py) ------------------------------------
bool parse_blob(const void* blob, size_t nbytes, std::map<std::string, HostTensor>& out, std::string& err) {
    const uint8_t* p = static_cast<const uint8_t*>(blob);
    if (nbytes < 16 || memcmp(p, "VAPW0001", 8) != 0) {
        err = "weight blob: bad magic";
        return false;
    }
    uint32_t n, tb;
    memcpy(&n, p + 8, 4);
    memcpy(&tb, p + 12, 4);
    if ((uint64_t)n * 96u != (uint64_t)tb || 16 + (uint64_t)tb > (uint64_t)nbytes) {
        err = "weight blob: corrupt table";
        return false;
    }
    for (uint32_t i = 0; i < n; ++i) {
        const uint8_t* e = p + 16 + (size_t)i * 96;
        char name[65];
        memcpy(name, e, 64);
        name[64] = 0;
        HostTensor t;
        uint32_t ndim;
        memcpy(&ndim, e + 64, 4);
        memcpy(t.dims, e + 68, 16);
        uint64_t off;
        uint32_t nb;
        memcpy(&off, e + 84, 8);
        memcpy(&nb, e + 92, 4);
        uint64_t prod = 1;
        for (uint32_t d = 0; d < ndim && d < 4; ++d) prod *= t.dims[d];
        if (ndim > 4 || off > (uint64_t)nbytes || (uint64_t)nb > (uint64_t)nbytes - off || (off & 3) || prod * 4 != (uint64_t)nb) {
            err = std::string("weight blob: bad entry ") + name;
            return false;
        }
        t.ndim = (int)ndim;
        t.numel = nb / 4;
        t.data = reinterpret_cast<const float*>(p + off);
        out[name] = t;
    }
    return true;
}

struct Loader {
    vapb_ctx* c;
    std::map<std::string, HostTensor> t;
    std::string err;
    bool ok = true;

    const HostTensor* get(const std::string& name, std::initializer_list<uint32_t> shape) {
        auto it = t.find(name);
        if (it == t.end()) {
            if (ok) err = "missing tensor " + name;
            ok = false;
            return nullptr;
        }
        const HostTensor& h = it->second;
        size_t numel = 1;
        int i = 0;
        bool match = (int)shape.size() == h.ndim;
        for (uint32_t d : shape) {
            if (match && h.dims[i] != d) match = false;
            numel *= d;
            ++i;
        }
        if (!match || numel != h.numel) {
            if (ok) err = "wrong shape for " + name;
            ok = false;
            return nullptr;
        }
        return &h;
    }
    float* upload(const std::vector<float>& v) {
        float* d = nullptr;
        if (cudaMalloc(&d, std::max<size_t>(v.size(), 4) * sizeof(float)) != cudaSuccess) {
            if (ok) err = "cudaMalloc failed (weights)";
            ok = false;
            return nullptr;
        }
        c->allocs.push_back(d);
        cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice);
        return d;
    }
    float* up(const std::string& name, std::initializer_list<uint32_t> shape) {
        const HostTensor* h = get(name, shape);
        if (!h) return nullptr;
        return upload(std::vector<float>(h->data, h->data + h->numel));
    }
    // [N][K] row-major -> [K][N] (k-major), for kernels that read a weight row slice per warp
    float* upT(const std::string& name, uint32_t N, uint32_t K) {
        const HostTensor* h = get(name, {N, K});
        if (!h) return nullptr;
        std::vector<float> v((size_t)N * K);
        for (uint32_t n = 0; n < N; ++n)
            for (uint32_t k = 0; k < K; ++k) v[(size_t)k * N + n] = h->data[(size_t)n * K + k];
        return upload(v);
    }
    // Conv1d weight [Cout][Cin][k] -> GEMM weight [Cout][k*Cin] (K index = tap*Cin + cin), matching
    // the channels-last activation rows (tap-major K).
    std::vector<float> conv_as_gemm(const std::string& name, uint32_t k) {
        const HostTensor* h = get(name, {256u, 256u, k});
        std::vector<float> v;
        if (!h) return v;
        v.resize((size_t)256 * k * 256);
        for (uint32_t n = 0; n < 256; ++n)
            for (uint32_t ci = 0; ci < 256; ++ci)
                for (uint32_t tap = 0; tap < k; ++tap)
                    v[(size_t)n * k * 256 + tap * 256 + ci] = h->data[((size_t)n * 256 + ci) * k + tap];
        return v;
    }
    std::vector<float> concat(std::initializer_list<const HostTensor*> parts) {
        std::vector<float> v;
        for (const HostTensor* h : parts)
            if (h) v.insert(v.end(), h->data, h->data + h->numel);
        return v;
    }
};

template <typename Tp>
int dalloc(vapb_ctx* c, Tp** p, size_t n, bool zero = true) {
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(Tp));
    if (e != cudaSuccess) return fail(c, VAPB_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", n * sizeof(Tp), cudaGetErrorString(e));
    c->allocs.push_back(d);
    if (zero) cudaMemset(d, 0, std::max<size_t>(n, 1) * sizeof(Tp));
    *p = static_cast<Tp*>(d);
    return 0;
}

RowMap act_map(const vapb_ctx* c, int layer) {      // logical row = chunk*L + position
    RowMap m;
    m.rpc = c->L[layer];
    m.chunk_stride = (long long)(c->L[layer] + 2 * c->halo[layer]) * kD;
    m.row_stride = kD;
    m.offset = (long long)c->halo[layer] * kD;
    return m;
}

int load_layer(Loader& ld, LayerWeights& lw, const std::string& p, bool cross) {
    lw.cross = cross;
    lw.ln_sa_w = ld.up(p + "ln_self_attn.weight", {256});
    lw.ln_sa_b = ld.up(p + "ln_self_attn.bias", {256});
    lw.ln_ff_w = ld.up(p + "ln_ffnetwork.weight", {256});
    lw.ln_ff_b = ld.up(p + "ln_ffnetwork.bias", {256});
    lw.sa.Wqkv = ld.upload(ld.concat({ld.get(p + "mha.query.weight", {256, 256}), ld.get(p + "mha.key.weight", {256, 256}),
                                      ld.get(p + "mha.value.weight", {256, 256})}));
    lw.sa.Wproj = ld.up(p + "mha.proj.weight", {256, 256});
    lw.sa.slopes = ld.up(p + "mha.m", {4});
    lw.W1 = ld.up(p + "ffnetwork.0.weight", {768, 256});
    lw.W2 = ld.up(p + "ffnetwork.3.weight", {256, 768});
    if (cross) {
        lw.ln_src_w = ld.up(p + "ln_src_attn.weight", {256});
        lw.ln_src_b = ld.up(p + "ln_src_attn.bias", {256});
        lw.Wq_c = ld.up(p + "mha_cross.query.weight", {256, 256});
        lw.Wkv_c = ld.upload(ld.concat({ld.get(p + "mha_cross.key.weight", {256, 256}),
                                        ld.get(p + "mha_cross.value.weight", {256, 256})}));
        lw.Wproj_c = ld.up(p + "mha_cross.proj.weight", {256, 256});
        lw.slopes_c = ld.up(p + "mha_cross.m", {4});
    }
    return 0;
}

// ---- the step ---------------------------------------------------------------------------
struct Step {
    vapb_ctx* c;
    cudaStream_t st;
    int B;
    const float* audio;
    float* out;
    const IoPtrs* io = nullptr;     // graph capture: kernels read audio / out through this device record instead
    int n = 0;      // launches
    std::vector<std::pair<const char*, cudaEvent_t>>* prof = nullptr;   // per-launch events (vapb_profile_step)
};

// bookkeeping after every kernel launch: count it and, when profiling, drop an event behind it
void mark(Step& s, const char* tag, int launches = 1) {
    s.n += launches;
    if (s.prof) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s.st);
        s.prof->push_back({tag, e});
    }
}

// C = act(A W^T + bias) + R through the selected GEMM engine.
void gemm(Step& s, const char* tag, const float* A, RowMap amap, const float* W, const TcWeight* tcw, const float* bias,
          const float* R, RowMap rmap, float* C, RowMap cmap, int M, int N, int K, int act, const float* ln_w = nullptr,
          const float* ln_b = nullptr, int ksplit = 1, long long csplit_stride = 0, int four_products = 0) {
    GemmArgs g;
    g.ln_w = ln_w; g.ln_b = ln_b;
    g.ksplit = ksplit; g.csplit_stride = csplit_stride;
    g.four_products = four_products;
    g.A = A; g.amap = amap; g.W = W; g.bias = bias; g.R = R; g.rmap = rmap; g.C = C; g.cmap = cmap;
    g.M = M; g.N = N; g.K = K; g.act = act;
    if (s.c->opt_gemm == 1 && tcw && tcw->hi) {
        mark(s, tag, launch_gemm_tc(g, *tcw, s.c->tcws, s.st));
    } else {
        launch_sgemm(g, s.st);
        mark(s, tag);
    }
}

void tap_copy(Step& s, float* dst, const float* src, size_t n) {
    if (s.c->opt_keep_taps && dst) cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s.st);
}

void attention(Step& s, const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, float* O,
               const float* slopes, int sibling, bool ring = false) {
    AttnArgs a;
    a.Q = Q; a.ldq = ldq; a.K = K; a.ldk = ldk; a.V = V; a.ldv = ldv; a.O = O; a.ldo = kD;
    a.tvalid = s.c->tvalid; a.slopes = slopes; a.n_seq = 2 * s.B; a.T = s.c->T; a.sibling = sibling;
    if (ring) { a.ring_ids = s.c->ids_dev; a.ring_count = s.c->count; }
    launch_attention(a, s.st);
    mark(s, sibling ? "attn_cross" : "attn_self");
}

bool qkv_cache_active(const vapb_ctx* c) { return c->opt_qkv_cache && c->opt_gemm == 1 && c->opt_fuse_ln && !c->opt_keep_taps && c->qkv_ring; }

void transformer_layer(Step& s, const LayerWeights& lw, bool qkv_cached = false) {
    vapb_ctx* c = s.c;
    const int R = 2 * s.B * c->T;
    const RowMap pd = plain_map(kD), pf = plain_map(kFF), p3 = plain_map(3 * kD), p2 = plain_map(2 * kD);
    // K/V of the cross attention come from the RAW layer input of the sibling channel
    // (modules.py:276-283): projected from X before the proj GEMM updates X in place.  With "fork" the
    // projection runs on a side branch, concurrently with LN -> QKV -> attention.
    const bool fork = lw.cross && c->opt_fork && s.prof == nullptr;
    if (lw.cross) {
        cudaStream_t main_st = s.st;
        if (fork) {
            cudaEventRecord(c->ev_fork, main_st);
            cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0);
            s.st = c->side_stream;
        }
        gemm(s, "gemm_kv_cross", c->X, pd, lw.Wkv_c, &lw.tc_kv_c, nullptr, nullptr, pd, c->KVc, p2, R, 2 * kD, kD, 0);
        if (fork) {
            cudaEventRecord(c->ev_join, c->side_stream);
            s.st = main_st;
        }
    }
    // self attention block (modules.py:268-272)
    // LayerNorm is a prologue of the consuming GEMM on the tensor-core path, a kernel of its own otherwise
    const bool fuse_ln = c->opt_gemm == 1 && c->opt_fuse_ln;
    if (qkv_cached) {
        // Q / K / V of every window frame are already in the stream's ring: only the newest frame is projected (2B rows)
        gemm(s, "gemm_ln_qkv_new", c->ebuf, pd, lw.sa.Wqkv, &lw.sa.tc_qkv, nullptr, nullptr, pd, c->QKVn, p3, 2 * s.B, 3 * kD, kD, 0, lw.ln_sa_w, lw.ln_sa_b);
        launch_qkv_append(c->QKVn, c->qkv_ring, c->count, c->ids_dev, s.B, c->T, s.st); mark(s, "qkv_append");
        attention(s, c->qkv_ring, 3 * kD, c->qkv_ring + kD, 3 * kD, c->qkv_ring + 2 * kD, 3 * kD, c->O, lw.sa.slopes, 0, true);
    } else {
    if (fuse_ln) {
        gemm(s, "gemm_ln_qkv", c->X, pd, lw.sa.Wqkv, &lw.sa.tc_qkv, nullptr, nullptr, pd, c->QKV, p3, R, 3 * kD, kD, 0, lw.ln_sa_w, lw.ln_sa_b);
    } else {
        launch_layernorm(c->X, pd, c->Z, pd, R, lw.ln_sa_w, lw.ln_sa_b, 0, s.st); mark(s, "layernorm");
        gemm(s, "gemm_qkv", c->Z, pd, lw.sa.Wqkv, &lw.sa.tc_qkv, nullptr, nullptr, pd, c->QKV, p3, R, 3 * kD, kD, 0);
    }
    attention(s, c->QKV, 3 * kD, c->QKV + kD, 3 * kD, c->QKV + 2 * kD, 3 * kD, c->O, lw.sa.slopes, 0);
    }
    if (fork) cudaStreamWaitEvent(s.st, c->ev_join, 0);      // X is overwritten next: the side branch must have read it
    gemm(s, "gemm_proj", c->O, pd, lw.sa.Wproj, &lw.sa.tc_proj, nullptr, c->X, pd, c->X, pd, R, kD, kD, 0);
    if (lw.cross) {
        if (fuse_ln) {
            gemm(s, "gemm_ln_q_cross", c->X, pd, lw.Wq_c, &lw.tc_q_c, nullptr, nullptr, pd, c->Qc, pd, R, kD, kD, 0, lw.ln_src_w, lw.ln_src_b);
        } else {
            launch_layernorm(c->X, pd, c->Z, pd, R, lw.ln_src_w, lw.ln_src_b, 0, s.st); mark(s, "layernorm");
            gemm(s, "gemm_q_cross", c->Z, pd, lw.Wq_c, &lw.tc_q_c, nullptr, nullptr, pd, c->Qc, pd, R, kD, kD, 0);
        }
        attention(s, c->Qc, kD, c->KVc, 2 * kD, c->KVc + kD, 2 * kD, c->O, lw.slopes_c, 1);
        gemm(s, "gemm_proj", c->O, pd, lw.Wproj_c, &lw.tc_proj_c, nullptr, c->X, pd, c->X, pd, R, kD, kD, 0);
    }
    // feed forward (modules.py:9-21, 285): Linear(256,768) -> GELU -> Linear(768,256), no biases
    if (fuse_ln) {
        gemm(s, "gemm_ln_ffn1", c->X, pd, lw.W1, &lw.tc_w1, nullptr, nullptr, pd, c->Hd, pf, R, kFF, kD, 1, lw.ln_ff_w, lw.ln_ff_b);
    } else {
        launch_layernorm(c->X, pd, c->Z, pd, R, lw.ln_ff_w, lw.ln_ff_b, 0, s.st); mark(s, "layernorm");
        gemm(s, "gemm_ffn1", c->Z, pd, lw.W1, &lw.tc_w1, nullptr, nullptr, pd, c->Hd, pf, R, kFF, kD, 1);
    }
    gemm(s, "gemm_ffn2", c->Hd, pf, lw.W2, &lw.tc_w2, nullptr, c->X, pd, c->X, pd, R, kD, kFF, 0);
}

// Final cross layer with the query side restricted to the newest frame of every sequence (exact:
// nothing downstream reads the other positions, vap_main.py:316-317).  K/V still cover the window.
void transformer_layer_last(Step& s, const LayerWeights& lw, bool kv_done = false, bool query_side = true) {
    vapb_ctx* c = s.c;
    const int NL = 2 * s.B, R = NL * c->T;
    const RowMap pd = plain_map(kD), pf = plain_map(kFF), p2 = plain_map(2 * kD);
    const bool fuse_ln = c->opt_gemm == 1 && c->opt_fuse_ln;
    auto ln_gemm = [&](const char* tag, const float* A, float* Z, int M, const float* lnw, const float* lnb, const float* W,
                       const TcWeight* tcw, float* C, RowMap cm, int N, int act) {
        if (fuse_ln) {
            gemm(s, tag, A, pd, W, tcw, nullptr, nullptr, pd, C, cm, M, N, kD, act, lnw, lnb);
        } else {
            launch_layernorm(A, pd, Z, pd, M, lnw, lnb, 0, s.st); mark(s, "layernorm");
            gemm(s, tag, Z, pd, W, tcw, nullptr, nullptr, pd, C, cm, M, N, kD, act);
        }
    };
    auto attn_last = [&](const float* Q, const float* K, const float* V, int ldkv, const float* slopes, int sibling) {
        AttnArgs a;
        a.Q = Q; a.ldq = kD; a.K = K; a.ldk = ldkv; a.V = V; a.ldv = ldkv; a.O = c->Ol; a.ldo = kD;
        a.tvalid = c->tvalid; a.slopes = slopes; a.n_seq = NL; a.T = c->T; a.sibling = sibling;
        launch_attention_last(a, s.st);
        mark(s, sibling ? "attn_cross_last" : "attn_self_last");
    };
    // keys / values over the whole window: cross attention from the RAW sibling input, self attention from LN(X)
    if (!kv_done) {
        gemm(s, "gemm_kv_cross", c->X, pd, lw.Wkv_c, &lw.tc_kv_c, nullptr, nullptr, pd, c->KVc, p2, R, 2 * kD, kD, 0);
        ln_gemm("gemm_ln_kv_self", c->X, c->Z, R, lw.ln_sa_w, lw.ln_sa_b, lw.sa.Wqkv + (size_t)kD * kD, &lw.sa.tc_kv, c->QKV, p2, 2 * kD, 0);
        launch_gather_last(c->X, c->tvalid, c->Xl, NL, c->T, s.st); mark(s, "gather_last");
    }
    if (!query_side) return;          // the fused tail kernel (k_tail) does the rest
    // self attention of the newest frame
    ln_gemm("gemm_ln_q_last", c->Xl, c->Zl, NL, lw.ln_sa_w, lw.ln_sa_b, lw.sa.Wqkv, &lw.sa.tc_q, c->Ql, pd, kD, 0);
    attn_last(c->Ql, c->QKV, c->QKV + kD, 2 * kD, lw.sa.slopes, 0);
    gemm(s, "gemm_proj_last", c->Ol, pd, lw.sa.Wproj, &lw.sa.tc_proj, nullptr, c->Xl, pd, c->Xl, pd, NL, kD, kD, 0);
    // cross attention
    ln_gemm("gemm_ln_q_cross_last", c->Xl, c->Zl, NL, lw.ln_src_w, lw.ln_src_b, lw.Wq_c, &lw.tc_q_c, c->Ql, pd, kD, 0);
    attn_last(c->Ql, c->KVc, c->KVc + kD, 2 * kD, lw.slopes_c, 1);
    gemm(s, "gemm_proj_last", c->Ol, pd, lw.Wproj_c, &lw.tc_proj_c, nullptr, c->Xl, pd, c->Xl, pd, NL, kD, kD, 0);
    // feed forward
    ln_gemm("gemm_ln_ffn1_last", c->Xl, c->Zl, NL, lw.ln_ff_w, lw.ln_ff_b, lw.W1, &lw.tc_w1, c->Hl, pf, kFF, 1);
    gemm(s, "gemm_ffn2_last", c->Hl, pf, lw.W2, &lw.tc_w2, nullptr, c->Xl, pd, c->Xl, pd, NL, kD, kFF, 0);
}

// ---- op list of the per-stream persistent transformer kernel (fused_tf.cu) ---------------
FOp fop_gemm(const float* A, int lda, int K, const TcWeight& w, const float* ln_w, const float* ln_b, const float* R, float* C,
             int ldc, int N, int act) {
    FOp o;
    memset(&o, 0, sizeof o);
    o.f.kind = FOP_GEMM;
    o.map_hi = w.map_hi[0];
    o.map_lo = w.map_lo[0];
    o.f.A = A; o.f.lda = lda; o.f.K = K; o.f.ln_w = ln_w; o.f.ln_b = ln_b; o.f.R = R; o.f.C = C; o.f.ldc = ldc; o.f.N = N; o.f.act = act;
    return o;
}
FOp fop_attn(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, float* O, int ldo, const float* slopes, int sibling) {
    FOp o;
    memset(&o, 0, sizeof o);
    o.f.kind = FOP_ATTN;
    o.f.Q = Q; o.f.ldq = ldq; o.f.Kp = K; o.f.ldk = ldk; o.f.V = V; o.f.ldv = ldv; o.f.O = O; o.f.ldo = ldo; o.f.slopes = slopes; o.f.sibling = sibling;
    return o;
}
FOp fop_misc(int kind) {
    FOp o;
    memset(&o, 0, sizeof o);
    o.f.kind = kind;
    return o;
}

int build_fused_ops(vapb_ctx* c) {
    std::vector<FOp> ops;
    ops.push_back(fop_misc(FOP_GATHER_RING));
    for (int l = 0; l < 3; ++l) {
        const LayerWeights& lw = c->layers[l];
        const size_t first = ops.size();
        if (lw.cross) ops.push_back(fop_gemm(c->X, kD, kD, lw.tc_kv_c, nullptr, nullptr, nullptr, c->KVc, 2 * kD, 2 * kD, 0));
        ops.push_back(fop_gemm(c->X, kD, kD, lw.sa.tc_qkv, lw.ln_sa_w, lw.ln_sa_b, nullptr, c->QKV, 3 * kD, 3 * kD, 0));
        // The [rows][768] Q | K | V buffer carries every other intermediate of the layer as well (a smaller footprint keeps the
        // scratch rows in L2 instead of streaming dead lines to HBM): the attention writes O over the Q columns of the heads
        // it has loaded, the cross attention's Q projection and O reuse those 256 columns once the self-attention's
        // projection has consumed them, and the FFN's hidden rows (768 wide, the same pitch) replace Q | K | V altogether.
        float* const Osc = c->QKV;
        ops.push_back(fop_attn(c->QKV, 3 * kD, c->QKV + kD, 3 * kD, c->QKV + 2 * kD, 3 * kD, Osc, 3 * kD, lw.sa.slopes, 0));
        ops.push_back(fop_gemm(Osc, 3 * kD, kD, lw.sa.tc_proj, nullptr, nullptr, c->X, c->X, kD, kD, 0));
        if (lw.cross) {
            ops.push_back(fop_gemm(c->X, kD, kD, lw.tc_q_c, lw.ln_src_w, lw.ln_src_b, nullptr, Osc, 3 * kD, kD, 0));
            ops.push_back(fop_attn(Osc, 3 * kD, c->KVc, 2 * kD, c->KVc + kD, 2 * kD, Osc, 3 * kD, lw.slopes_c, 1));
            ops.push_back(fop_gemm(Osc, 3 * kD, kD, lw.tc_proj_c, nullptr, nullptr, c->X, c->X, kD, kD, 0));
        }
        static_assert(kFF == 3 * kD, "the FFN hidden rows alias the Q | K | V rows");
        ops.push_back(fop_gemm(c->X, kD, kD, lw.tc_w1, lw.ln_ff_w, lw.ln_ff_b, nullptr, c->QKV, kFF, kFF, 1));
        ops.push_back(fop_gemm(c->QKV, kFF, kFF, lw.tc_w2, nullptr, nullptr, c->X, c->X, kD, kD, 0));
        // vad reads the ar_channel output: a side task of the first op of cross layer 0 (which only reads X)
        if (l == 1 && c->head_kind == VAPB_HEAD_VAP) ops[first].f.side = FSIDE_VAD;
    }
    {   // pruned last layer: only its window-wide K / V projections and the newest-frame gather
        const LayerWeights& lw = c->layers[3];
        ops.push_back(fop_gemm(c->X, kD, kD, lw.tc_kv_c, nullptr, nullptr, nullptr, c->KVc, 2 * kD, 2 * kD, 0));
        ops.back().f.side = FSIDE_GATHER_LAST;
        ops.push_back(fop_gemm(c->X, kD, kD, lw.sa.tc_kv, lw.ln_sa_w, lw.ln_sa_b, nullptr, c->QKV, 2 * kD, 2 * kD, 0));
    }
    c->n_fops = (int)ops.size();
    void* d = nullptr;
    if (cudaMalloc(&d, ops.size() * sizeof(FOp)) != cudaSuccess) return fail(c, VAPB_ENOMEM, "cudaMalloc(fused op list) failed");
    c->allocs.push_back(d);
    if (cudaMemcpy(d, ops.data(), ops.size() * sizeof(FOp), cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(c, VAPB_ECUDA, "cudaMemcpy(fused op list) failed");
    c->fops = static_cast<FOp*>(d);
    if (cudaMalloc(&d, 64 * sizeof(long long)) != cudaSuccess) return fail(c, VAPB_ENOMEM, "cudaMalloc(fused clocks) failed");
    c->allocs.push_back(d);
    cudaMemset(d, 0, 64 * sizeof(long long));
    c->fused_clk = static_cast<long long*>(d);
    std::string err;
    if (!fused_prepare(err)) return fail(c, VAPB_ECUDA, "%s", err.c_str());
    return 0;
}


// ---- stream kernel v2: LayerNorm folded into the weights (fused_tf2.cuh) -------------------
// rows of `out` += W[n][k] * g[k] (g = LayerNorm gain or null); s[n] = sum_k W'[n][k]; c[n] = sum_k b[k] W[n][k]
void v2_append(std::vector<float>& out, std::vector<float>& sv, std::vector<float>& cv, const HostTensor* W, const HostTensor* g,
               const HostTensor* b) {
    if (!W) return;
    const size_t N = W->numel / kD;
    for (size_t n = 0; n < N; ++n) {
        double ssum = 0.0, csum = 0.0;
        for (int k = 0; k < kD; ++k) {
            const float w = W->data[n * kD + k];
            const float ws = g ? w * g->data[k] : w;
            out.push_back(ws);
            ssum += ws;
            if (b) csum += (double)b->data[k] * w;
        }
        if (g) {
            sv.push_back((float)ssum);
            cv.push_back((float)csum);
        }
    }
}

void build_v2_weights(Loader& ld, vapb_ctx* c) {
    for (int l = 0; l < 4; ++l) {
        const std::string p = l == 0 ? std::string("ar_channel.layers.0.") : "ar.layers." + std::to_string(l - 1) + ".";
        vapb_ctx::V2Layer& v = c->v2[l];
        const HostTensor* g_sa = ld.get(p + "ln_self_attn.weight", {256});
        const HostTensor* b_sa = ld.get(p + "ln_self_attn.bias", {256});
        std::vector<float> W, sv, cv;
        if (l < 3) v2_append(W, sv, cv, ld.get(p + "mha.query.weight", {256, 256}), g_sa, b_sa);
        v2_append(W, sv, cv, ld.get(p + "mha.key.weight", {256, 256}), g_sa, b_sa);
        v2_append(W, sv, cv, ld.get(p + "mha.value.weight", {256, 256}), g_sa, b_sa);
        v.nln_g1 = (int)sv.size();
        if (l > 0) {
            v2_append(W, sv, cv, ld.get(p + "mha_cross.key.weight", {256, 256}), nullptr, nullptr);
            v2_append(W, sv, cv, ld.get(p + "mha_cross.value.weight", {256, 256}), nullptr, nullptr);
        }
        v.n_g1 = (int)(W.size() / kD);
        v.Wg1 = ld.upload(W);
        v.s_g1 = ld.upload(sv);
        v.c_g1 = ld.upload(cv);
        if (l > 0 && l < 3) {
            std::vector<float> Wq, sq, cq;
            v2_append(Wq, sq, cq, ld.get(p + "mha_cross.query.weight", {256, 256}), ld.get(p + "ln_src_attn.weight", {256}),
                      ld.get(p + "ln_src_attn.bias", {256}));
            v.Wqc = ld.upload(Wq);
            v.s_qc = ld.upload(sq);
            v.c_qc = ld.upload(cq);
        }
        if (l < 3) {
            std::vector<float> W1, s1, c1;
            v2_append(W1, s1, c1, ld.get(p + "ffnetwork.0.weight", {768, 256}), ld.get(p + "ln_ffnetwork.weight", {256}),
                      ld.get(p + "ln_ffnetwork.bias", {256}));
            v.W1s = ld.upload(W1);
            v.s_1 = ld.upload(s1);
            v.c_1 = ld.upload(c1);
        }
    }
}

int build_fused2(vapb_ctx* c) {
    if (c->T > 64) return 0;                       // v2 covers the both-channels-in-one-tile geometry only
    std::string err;
    for (int l = 0; l < 4; ++l) {
        vapb_ctx::V2Layer& v = c->v2[l];
        bool ok = tc_prepare_weight(v.Wg1, v.n_g1, kD, v.g1, c->allocs, err);
        if (ok && v.Wqc) ok = tc_prepare_weight(v.Wqc, kD, kD, v.qc, c->allocs, err);
        if (ok && v.W1s) ok = tc_prepare_weight(v.W1s, kFF, kD, v.w1, c->allocs, err);
        if (!ok) return fail(c, VAPB_ECUDA, "stream kernel v2 weights: %s", err.c_str());
    }
    const size_t R2 = ((size_t)c->max_batch + 1) * 128;      // + one scratch slot for the ghost stream of an odd batch
    int rc = 0;
#define DA2(ptr, n) if (!rc) rc = dalloc(c, &(ptr), (n))
    DA2(c->X2f, R2 * kD);
    DA2(c->St2, R2 * 16);
    DA2(c->X2h, R2 * kD);
    DA2(c->X2l, R2 * kD);
    DA2(c->G1h, R2 * 1280);
    DA2(c->G1l, R2 * 1280);
#undef DA2
    if (rc) return rc;
    // box rows 128 / 64 (loads: 64 columns, SWIZZLE_128B) and 32 (epilogue stores: 32-column half boxes, SWIZZLE_64B, over
    // [sequence][position < T][cols]: the padding rows T..63 of a sequence are clipped, never written, and stay zero)
    struct Planes { CUtensorMap hi128, lo128, hi64, lo64, hi32h, lo32h; };
    auto planes = [&](__nv_bfloat16* hi, __nv_bfloat16* lo, size_t cols, Planes& m) {
        return tc_encode_bf16_2d(&m.hi128, hi, R2, cols, 128, err) && tc_encode_bf16_2d(&m.lo128, lo, R2, cols, 128, err) &&
               tc_encode_bf16_2d(&m.hi64, hi, R2, cols, 64, err) && tc_encode_bf16_2d(&m.lo64, lo, R2, cols, 64, err) &&
               tc_encode_bf16_3d_half(&m.hi32h, hi, R2 / 64, c->T, cols, 64, err) && tc_encode_bf16_3d_half(&m.lo32h, lo, R2 / 64, c->T, cols, 64, err);
    };
    // ONE plane pair of 1 280 columns per row carries every intermediate of a layer (less footprint in L2 = fewer dead lines
    // written back to HBM): [Q | K | V | K cross | V cross] of the fused projection; the attention writes O over the Q columns
    // of its own heads (Q is in shared memory by then), the cross attention's Q projection and its O go to the same 256
    // columns once the self-attention's projection has consumed them, and the FFN's hidden rows take columns 0..767
    // (Q / K / V of the self attention are dead; the cross K / V columns are never touched).
    Planes mX, mG1;
    if (!planes(c->X2h, c->X2l, kD, mX) || !planes(c->G1h, c->G1l, 1280, mG1))
        return fail(c, VAPB_ECUDA, "stream kernel v2 tensor maps: %s", err.c_str());
    const Planes &mO = mG1, &mQc = mG1, &mH = mG1;
    CUtensorMap mXf, mKVs, mKVc;           // fp32 store targets: residual stream; K | V rows of the pruned layer for the tail
    if (!tc_encode_f32_3d(&mXf, c->X2f, R2 / 64, c->T, kD, 32, err, 64) || !tc_encode_f32_3d(&mKVs, c->QKV, (size_t)2 * c->max_batch, c->T, 512, 32, err) ||
        !tc_encode_f32_3d(&mKVc, c->KVc, (size_t)2 * c->max_batch, c->T, 512, 32, err))
        return fail(c, VAPB_ECUDA, "stream kernel v2 tensor maps: %s", err.c_str());

    std::vector<F2Op> ops;
    auto gemm = [&](const Planes& A, int K, const TcWeight& w, int N, int out_mode, int n_ln, const float* ls, const float* lc, int act,
                    const Planes* out, __nv_bfloat16* oh, __nv_bfloat16* ol, int ld_out, int cta_sync) {
        F2Op o;
        memset(&o, 0, sizeof o);
        o.m[0] = A.hi64; o.m[1] = A.lo64; o.m[2] = w.map_hi[0]; o.m[3] = w.map_lo[0];
        if (out_mode == F2_OUT_X) { o.m[4] = mX.hi32h; o.m[5] = mX.lo32h; o.m[6] = mXf; }
        else if (out_mode == F2_OUT_F32) { o.m[4] = mKVs; o.m[5] = mKVc; }
        else { o.m[4] = out->hi32h; o.m[5] = out->lo32h; }
        o.f.kind = F2_GEMM; o.f.K = K; o.f.N = N; o.f.out_mode = out_mode; o.f.n_ln = n_ln; o.f.ln_s = ls; o.f.ln_c = lc; o.f.act = act;
        o.f.out_hi = oh; o.f.out_lo = ol; o.f.ld_out = ld_out; o.f.cta_sync = cta_sync;
        ops.push_back(o);
    };
    auto attn = [&](const Planes& Q, int qcol, int kcol, int vcol, const float* slopes, int sibling) {
        F2Op o;
        memset(&o, 0, sizeof o);
        o.m[0] = Q.hi128; o.m[1] = Q.lo128; o.m[2] = mG1.hi64; o.m[3] = mG1.lo64; o.m[4] = mO.hi32h; o.m[5] = mO.lo32h;
        o.f.kind = F2_ATTN; o.f.qcol = qcol; o.f.kcol = kcol; o.f.vcol = vcol; o.f.slopes = slopes; o.f.sibling = sibling;
        o.f.out_hi = c->G1h; o.f.out_lo = c->G1l; o.f.ld_out = 1280;
        ops.push_back(o);
    };
    {
        F2Op o;
        memset(&o, 0, sizeof o);
        o.f.kind = F2_GATHER;
        ops.push_back(o);
    }
    for (int l = 0; l < 3; ++l) {
        const LayerWeights& lw = c->layers[l];
        const vapb_ctx::V2Layer& v = c->v2[l];
        // Q / K / V of the self attention (LayerNorm folded) [+ K / V of the cross attention from the raw rows]: one op
        gemm(mX, kD, v.g1, v.n_g1, F2_OUT_PLANES, v.nln_g1, v.s_g1, v.c_g1, 0, &mG1, c->G1h, c->G1l, 1280, 1);
        if (l == 1 && c->head_kind == VAPB_HEAD_VAP) ops.back().f.side = F2_SIDE_VAD;      // reads the ar_channel output
        attn(mG1, 0, 256, 512, lw.sa.slopes, 0);
        gemm(mO, kD, lw.sa.tc_proj, kD, F2_OUT_X, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, kD, 0);
        if (lw.cross) {
            gemm(mX, kD, v.qc, kD, F2_OUT_PLANES, kD, v.s_qc, v.c_qc, 0, &mQc, c->G1h, c->G1l, 1280, 1);
            attn(mQc, 0, 768, 1024, lw.slopes_c, 1);
            gemm(mO, kD, lw.tc_proj_c, kD, F2_OUT_X, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, kD, 0);
        }
        gemm(mX, kD, v.w1, kFF, F2_OUT_PLANES, kFF, v.s_1, v.c_1, 1, &mH, c->G1h, c->G1l, 1280, 0);
        gemm(mH, kFF, lw.tc_w2, kD, F2_OUT_X, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, kD, 0);
    }
    {   // pruned last layer: window-wide K / V (self: LayerNorm folded, cross: raw rows) as fp32 rows for the newest-frame tail
        const vapb_ctx::V2Layer& v = c->v2[3];
        gemm(mX, kD, v.g1, v.n_g1, F2_OUT_F32, v.nln_g1, v.s_g1, v.c_g1, 0, nullptr, nullptr, nullptr, 512, 0);
        ops.back().f.out_f = c->QKV;
        ops.back().f.out_f2 = c->KVc;
        ops.back().f.side = F2_SIDE_GATHER_LAST;
    }
    for (const F2Op& o : ops)
        if (o.f.kind == F2_GEMM && ((o.f.N & 255) || (o.f.K != 256 && o.f.K != 768))) return fail(c, VAPB_EINVAL, "stream kernel v2: bad op shape");
    c->n_f2ops = (int)ops.size();
    void* d = nullptr;
    if (cudaMalloc(&d, ops.size() * sizeof(F2Op)) != cudaSuccess) return fail(c, VAPB_ENOMEM, "cudaMalloc(v2 op list) failed");
    c->allocs.push_back(d);
    if (cudaMemcpy(d, ops.data(), ops.size() * sizeof(F2Op), cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(c, VAPB_ECUDA, "cudaMemcpy(v2 op list) failed");
    c->f2ops = static_cast<F2Op*>(d);
    if (!fused2_prepare(err)) return fail(c, VAPB_ECUDA, "%s", err.c_str());
    return 0;
}

void fused_transformer2(Step& s) {
    vapb_ctx* c = s.c;
    Fused2Params p;
    memset(&p, 0, sizeof p);
    p.ops = c->f2ops; p.n_ops = c->n_f2ops; p.T = c->T; p.B = s.B;
    p.ring = c->ring; p.ring_w = c->ring; p.count = c->count; p.ids = c->ids_dev; p.tvalid = c->tvalid;
    p.ds_part = c->fused_ds_part; p.ds_stride = c->fused_ds_stride; p.ds_nsplit = c->fused_ds_nsplit;
    p.ds_lnw = c->ds_lnw; p.ds_lnb = c->ds_lnb; p.e_out = c->ebuf;
    p.Xf = c->X2f; p.Xh = c->X2h; p.Xl = c->X2l; p.stats = c->St2; p.Xlast = c->Xl;
    p.va_w = c->va_w; p.va_b = c->va_b; p.out = s.out; p.io = s.io;
    p.dbg = c->opt_fused_dbg ? c->fused_clk : nullptr;
    p.dbg_op = c->opt_fused_dbg - 1;
    launch_fused_tf2(p, s.B, s.st);
    mark(s, "fused_tf");
}

void fused_transformer(Step& s) {
    vapb_ctx* c = s.c;
    FusedParams p;
    p.ops = c->fops; p.n_ops = c->n_fops; p.T = c->T; p.mode = (2 * c->T <= 128) ? 0 : 1;
    p.ring = c->ring; p.count = c->count; p.ids = c->ids_dev; p.tvalid = c->tvalid; p.X = c->X; p.Xl = c->Xl;
    p.va_w = c->va_w; p.va_b = c->va_b; p.out = s.out; p.io = s.io;
    p.ds_part = c->fused_ds_part; p.ds_stride = c->fused_ds_stride; p.ds_nsplit = c->fused_ds_nsplit;
    p.ds_lnw = c->ds_lnw; p.ds_lnb = c->ds_lnb; p.e_out = c->ebuf;
    p.dbg = c->opt_fused_dbg ? c->fused_clk : nullptr;
    p.dbg_op = c->opt_fused_dbg - 1;          // fused_dbg = 1 + index of the op that gets fine stamps
    launch_fused_tf(p, s.B, s.st);
    mark(s, "fused_tf");
}

// ---- ar_channel + the three cross layers over X (window rows, oldest first) with the batched per-op kernels ----
void transformer_batched(Step& s, bool prune, bool qkv_cached) {
    vapb_ctx* c = s.c;
    const int B = s.B, NC = 2 * B, T = c->T;
    cudaStream_t st = s.st;
    const size_t RX = (size_t)NC * T * kD;
    tap_copy(s, c->tap_xin, c->X, RX);
    // ---- ar_channel: one TransformerLayer per channel, shared weights (vap_main.py:285-286)
    transformer_layer(s, c->layers[0], qkv_cached);
    tap_copy(s, c->tap_chan, c->X, RX);
    if (c->head_kind == VAPB_HEAD_VAP) {
        launch_vad(c->X, c->tvalid, c->va_w, c->va_b, s.out, s.io, B, T, st); mark(s, "vad");
    }
    // ---- ar: three TransformerStereoLayers (modules.py:289-300, 395-423)
    for (int li = 0; li < 3; ++li) {
        if (li == 2 && prune) transformer_layer_last(s, c->layers[3], false, !c->opt_tail);
        else transformer_layer(s, c->layers[1 + li]);
        tap_copy(s, c->tap_cross[li], c->X, RX);
    }
}

// ---- newest-frame side of the pruned layer (k_tail) or combinator + head; advances `count` of the batch's streams ----
void tail_or_head(Step& s, bool prune, int* count) {
    vapb_ctx* c = s.c;
    const int B = s.B, T = c->T;
    cudaStream_t st = s.st;
    if (prune && c->opt_tail) {
        const LayerWeights& lw = c->layers[3];
        TailArgs t;
        t.Xl = c->Xl; t.KVs = c->QKV; t.KVc = c->KVc; t.tvalid = c->tvalid;
        t.WqT = c->t_WqT; t.WprojT = c->t_WprojT; t.WqcT = c->t_WqcT; t.WprojcT = c->t_WprojcT; t.W1T = c->t_W1T; t.W2T = c->t_W2T;
        t.WaT = c->t_WaT; t.WbT = c->t_WbT; t.WhT = c->t_WhT;
        t.ln_sa_w = lw.ln_sa_w; t.ln_sa_b = lw.ln_sa_b; t.ln_src_w = lw.ln_src_w; t.ln_src_b = lw.ln_src_b; t.ln_ff_w = lw.ln_ff_w; t.ln_ff_b = lw.ln_ff_b;
        t.comb_lnw = c->comb_lnw; t.comb_lnb = c->comb_lnb; t.bh = c->bh; t.slopes_s = lw.sa.slopes; t.slopes_c = lw.slopes_c;
        t.n_out = c->n_out; t.head_kind = c->head_kind; t.B = B; t.T = T;
        t.out = s.out; t.io = s.io; t.count = count; t.ids = c->ids_dev;
        launch_tail(t, st); mark(s, "tail");
        return;
    }
    HeadArgs h;
    h.X = prune ? c->Xl : c->X; h.compact = prune ? 1 : 0; h.tvalid = c->tvalid; h.Wa = c->Wa; h.Wb = c->Wb; h.lnw = c->comb_lnw; h.lnb = c->comb_lnb;
    h.Wh = c->Wh; h.bh = c->bh; h.n_out = c->n_out; h.out = s.out; h.io = s.io;
    h.comb_tap = c->opt_keep_taps ? c->tap_comb : nullptr;
    h.logits_tap = c->opt_keep_taps ? c->tap_logits : nullptr;
    h.count = count; h.ids = c->ids_dev; h.B = B; h.T = T; h.head_kind = c->head_kind;
    launch_head(h, st); mark(s, "head");
}

// conv0 .. conv4 on NC chunk rows of s.audio, then the LSTM input projection of the inner frames -> c->Gx
void encoder_convs(Step& s, int NC) {
    vapb_ctx* c = s.c;
    cudaStream_t st = s.st;
    // ---- CPC encoder on the newest chunk (encoder.py:58-80)
    launch_conv0(s.audio, s.io, NC, c->S, c->L[0], c->w0, c->b0, c->cn0w, c->cn0b, c->act[0], act_map(c, 0), st); mark(s, "conv0_cn_relu");
    for (int i = 0; i < 4; ++i) {
        const ConvLayer& cv = c->conv[i];
        // im2col-free: output row p of a chunk reads k*256 contiguous floats starting at input
        // row (s*p - pad); the zero halo rows of the input buffer are the conv padding.
        RowMap am;
        am.rpc = cv.Lout;
        am.chunk_stride = (long long)(cv.Lin + 2 * c->halo[i]) * kD;
        am.row_stride = (long long)cv.s * kD;
        am.offset = (long long)(c->halo[i] - cv.p) * kD;
        const RowMap cm = act_map(c, i + 1);
        const int Mc = NC * cv.Lout;
        // split-K factors are fixed per layer (conv3: 2, conv4: 4) so that a stream's arithmetic does not depend
        // on the batch size; shorter accumulation chains + fp32 summation also measurably help parity
        const int ks_layer[4] = {c->opt_conv12_ks & 0xf ? (c->opt_conv12_ks & 0xf) : 1, (c->opt_conv12_ks >> 4) ? (c->opt_conv12_ks >> 4) : 1, 2, 4};
        const int ks = (c->opt_gemm == 1 && c->opt_splitk && (size_t)Mc * kD <= c->part_stride) ? ks_layer[i] : 1;
        if (ks > 1) {
            // small M, long K: K is split over CTAs; ChannelNorm sums the partial products
            gemm(s, "gemm_conv_splitk", c->act[i], am, cv.W, &cv.tc, cv.b, nullptr, cm, c->part, plain_map(kD), Mc, kD, cv.k * kD, 0,
                 nullptr, nullptr, ks, (long long)c->part_stride, c->opt_conv4p);
            launch_cn_relu(c->act[i + 1], cm, Mc, cv.cnw, cv.cnb, st, c->part, ks, (long long)c->part_stride); mark(s, "cn_relu");
        } else {
            gemm(s, i == 0 ? "gemm_conv1" : "gemm_conv2_4", c->act[i], am, cv.W, &cv.tc, cv.b, nullptr, cm, c->act[i + 1], cm, Mc, kD, cv.k * kD, 0,
                 nullptr, nullptr, 1, 0, c->opt_conv4p);
            launch_cn_relu(c->act[i + 1], cm, Mc, cv.cnw, cv.cnb, st); mark(s, "cn_relu");
        }
    }
    // ---- LSTM over the inner frames z[:, 1:-1] with persistent (h, c) (encoder.py:76-77)
    {
        RowMap am;
        am.rpc = c->n_lstm;
        am.chunk_stride = (long long)(c->L[4] + 2 * c->halo[4]) * kD;
        am.row_stride = kD;
        am.offset = (long long)(c->halo[4] + 1) * kD;
        const RowMap g4 = plain_map(4 * kD);
        // input projection for all n_lstm frames at once (fp32), then the fused recurrence
        gemm(s, "gemm_lstm_x", c->act[4], am, c->Wih, c->opt_lstm_x_tc ? &c->tc_ih : nullptr, c->b_lstm, nullptr, g4, c->Gx, g4, NC * c->n_lstm, 4 * kD, kD, 0);
    }
}

void transformer_batched(Step& s, bool prune, bool qkv_cached);
void tail_or_head(Step& s, bool prune, int* count);
bool step_uses_stream(const vapb_ctx* c, int B);

void enqueue_step(Step& s) {
    vapb_ctx* c = s.c;
    vapb::g_use_pdl = c->opt_pdl != 0 && s.prof == nullptr;    // per-kernel event timing needs plain launches
    vapb::g_attn_rk = c->opt_attn_rk != 0;
    const int B = s.B, NC = 2 * B, T = c->T;
    cudaStream_t st = s.st;
    // side branch: pull the transformer / tail weights into L2 while the encoder runs
    const bool prefetch = c->opt_prefetch && c->opt_gemm == 1 && c->pf_n > 0 && s.prof == nullptr;
    if (prefetch) {
        cudaEventRecord(c->ev_fork, st);
        cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0);
        launch_l2_prefetch(c->pf_ptrs, c->pf_bytes, c->pf_n, c->pf_lines, c->side_stream);
        cudaEventRecord(c->ev_join, c->side_stream);
        s.n += 1;
    }
    encoder_convs(s, NC);
    if (prefetch) cudaStreamWaitEvent(st, c->ev_join, 0);      // joined early: the warm-up itself takes a few microseconds
    {
        const RowMap g4 = plain_map(4 * kD);
        if (c->opt_lstm_fused) {
            launch_lstm_recurrent(c->Gx, c->Whh, c->hS, c->cS, c->ids_dev, c->Y, NC, c->n_lstm, st); mark(s, "lstm_recurrent");
        } else {
            launch_gather_state(c->hS, c->cS, c->ids_dev, c->hW, c->cW, B, st); mark(s, "lstm_state");
            for (int t = 0; t < c->n_lstm; ++t) {
                RowMap rm;
                rm.rpc = 1;
                rm.chunk_stride = (long long)c->n_lstm * 4 * kD;
                rm.row_stride = 0;
                rm.offset = (long long)t * 4 * kD;
                gemm(s, "gemm_lstm_h", c->hW, plain_map(kD), c->Whh, nullptr, nullptr, c->Gx, rm, c->Gt, g4, NC, 4 * kD, kD, 0);
                launch_lstm_cell(c->Gt, c->hW, c->cW, c->Y, NC, c->n_lstm, t, st); mark(s, "lstm_cell");
            }
            launch_scatter_state(c->hS, c->cS, c->ids_dev, c->hW, c->cW, B, st); mark(s, "lstm_state");
        }
    }
    // per-stream cluster kernel or batched per-op kernels: see step_uses_stream
    const bool prune = c->opt_prune && !c->opt_keep_taps;      // taps want every position of every layer
    const bool v2 = c->opt_fused_v == 2 && c->f2ops;
    const bool use_stream = step_uses_stream(c, B);
    // ---- downsample conv over exactly n_lstm frames + LayerNorm + GELU -> ring
    {
        const int Kd = c->n_lstm * kD;
        const int ks = (c->opt_gemm == 1 && c->opt_splitk) ? (Kd / 64) / 2 : 1;      // 2 k-blocks per slice, any batch
        if (ks > 1) {
            const long long dstride = (long long)NC * kD;
            gemm(s, "gemm_downsample", c->Y, plain_map(Kd), c->Wds, &c->tc_ds, c->bds, nullptr, plain_map(kD), c->part, plain_map(kD), NC, kD, Kd, 0,
                 nullptr, nullptr, ks, dstride, c->opt_conv4p);
            if (use_stream) { c->fused_ds_part = c->part; c->fused_ds_stride = dstride; c->fused_ds_nsplit = ks; }
            else launch_ln_gelu_ring(c->part, B, c->ds_lnw, c->ds_lnb, c->ring, c->count, c->ids_dev, T, c->ebuf, st, ks, dstride);
        } else {
            gemm(s, "gemm_downsample", c->Y, plain_map(Kd), c->Wds, &c->tc_ds, c->bds, nullptr, plain_map(kD), c->dsout, plain_map(kD), NC, kD, Kd, 0);
            if (use_stream) { c->fused_ds_part = c->dsout; c->fused_ds_stride = 0; c->fused_ds_nsplit = 1; }
            else launch_ln_gelu_ring(c->dsout, B, c->ds_lnw, c->ds_lnb, c->ring, c->count, c->ids_dev, T, c->ebuf, st);
        }
        if (!use_stream) mark(s, "ln_gelu_ring");      // the stream kernel's gather op finishes the embedding otherwise
    }
    if (use_stream) {
        // ---- ring gather, ar_channel, vad, cross layers 0-1 and the K/V of the pruned last layer: ONE launch,
        //      a cluster of two CTAs per stream (fused_tf.cu); then the newest-frame tail of the last layer
        if (v2) fused_transformer2(s);
        else fused_transformer(s);
        transformer_layer_last(s, c->layers[3], true, !c->opt_tail);
    } else {
        // ---- window of the last T embeddings, oldest first (vap_main.py:274-283)
        launch_gather_ring(c->ring, c->count, c->ids_dev, c->X, c->tvalid, B, T, st); mark(s, "gather_ring");
        transformer_batched(s, prune, qkv_cache_active(c));
    }
    tail_or_head(s, prune, c->count);
}

// Does a batch of B run the per-stream cluster kernel (which recomputes layer 0) or the batched per-op kernels?
// One wave of clusters holds `cap` streams (74 on a 148-SM part, either generation).  Up to one wave the stream kernel
// always wins.  Beyond it the launch runs in waves whose last one costs as much as a full one, so it only pays when the
// waves are well filled: measured (tools/big_batch_paths.py, profiles/r02_q_*) T = 50: B = 128 / 192 +3 % / +6 %,
// B = 96 / 256 / 1024 -4 % / -5 % / -13 %; T = 100 (one channel per CTA, first generation): +12 ... +18 % at B = 128 / 256 / 1024.
bool step_uses_stream(const vapb_ctx* c, int B) {
    const bool prune = c->opt_prune && !c->opt_keep_taps;
    if (!(c->opt_gemm == 1 && c->opt_fused && prune && c->fops)) return false;
    if (c->opt_fused == 2) return true;
    const bool v2 = c->opt_fused_v == 2 && c->f2ops;
    const int cap = v2 ? 2 * (c->sm_count / 4) : c->sm_count / 2;
    if (B <= cap) return true;
    const int waves = (B + cap - 1) / cap;
    const bool filled = (long long)B * 100 >= 85ll * waves * cap;
    return c->T > 64 ? filled : (filled && waves <= 3);
}

// Layer-0 Q/K/V cache bookkeeping in front of a step: a stream whose cached rows do not cover all its frames (it was
// stepped by the cluster kernel, imported, or the cache was off) gets its whole window re-projected from the embedding
// ring: one LN + GEMM over its 2T ring rows, outside the step graph.  Steady state: nothing to do.
int qkv_cache_pre_step(vapb_ctx* c, const int* ids, int B, cudaStream_t st) {
    if (!qkv_cache_active(c) || step_uses_stream(c, B)) return 0;
    const LayerWeights& lw = c->layers[0];
    for (int i = 0; i < B; ++i) {
        const int id = ids[i];
        if (c->h_cnt[id] == 0 || c->h_qkv[id] == c->h_cnt[id]) continue;
        GemmArgs g;
        g.A = c->ring + (size_t)id * 2 * c->T * kD; g.amap = plain_map(kD);
        g.W = lw.sa.Wqkv; g.bias = nullptr; g.R = nullptr; g.rmap = plain_map(kD);
        g.C = c->qkv_ring + (size_t)id * 2 * c->T * 3 * kD; g.cmap = plain_map(3 * kD);
        g.M = 2 * c->T; g.N = 3 * kD; g.K = kD; g.act = 0; g.ln_w = lw.ln_sa_w; g.ln_b = lw.ln_sa_b;
        launch_gemm_tc(g, lw.sa.tc_qkv, c->tcws, st);
        c->h_qkv[id] = c->h_cnt[id];
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(c, VAPB_ECUDA, "Q/K/V cache rebuild failed: %s", cudaGetErrorString(e));
    return 0;
}
void qkv_cache_post_step(vapb_ctx* c, const int* ids, int B) {
    const bool cached = qkv_cache_active(c) && !step_uses_stream(c, B);
    for (int i = 0; i < B; ++i) {
        const int id = ids[i];
        const bool was_current = c->h_qkv[id] == c->h_cnt[id];
        c->h_cnt[id] += 1;
        if (cached && was_current) c->h_qkv[id] = c->h_cnt[id];
    }
}

int check_ids(vapb_ctx* c, const int* ids, int B) {
    if (B <= 0 || B > c->max_batch) return fail(c, VAPB_EINVAL, "B=%d outside 1..%d", B, c->max_batch);
    if (!ids) return fail(c, VAPB_EINVAL, "stream_ids is NULL");
    std::vector<int> v(ids, ids + B);
    std::sort(v.begin(), v.end());
    if (v.front() < 0 || v.back() >= c->max_streams) return fail(c, VAPB_EINVAL, "stream id outside 0..%d", c->max_streams - 1);
    for (int i = 1; i < B; ++i)
        if (v[i] == v[i - 1]) return fail(c, VAPB_EINVAL, "duplicate stream id %d in one batch", v[i]);
    return 0;
}

}  // namespace

// =========================================================================================
extern "C" {

const char* vapb_version(void) { return "vapb200 0.1 sm_100a"; }

const char* vapb_last_error(vapb_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int vapb_create(const void* weights_blob, size_t nbytes, int frame_hz, int ctx_frames, int max_streams,
                int max_batch, int head_kind, int device, vapb_handle* out) {
    if (!out) return fail(nullptr, VAPB_EINVAL, "out is NULL");
    *out = nullptr;
    if (!weights_blob) return fail(nullptr, VAPB_EINVAL, "weights_blob is NULL");
    if (frame_hz != 20 && frame_hz != 10 && frame_hz != 5)
        return fail(nullptr, VAPB_EUNSUPPORTED, "frame_hz %d not in {20,10,5}", frame_hz);
    if (ctx_frames < 1 || ctx_frames > kMaxT)
        return fail(nullptr, VAPB_EUNSUPPORTED, "ctx_frames %d outside 1..%d", ctx_frames, kMaxT);
    if (max_streams < 1 || max_batch < 1 || max_batch > max_streams)
        return fail(nullptr, VAPB_EINVAL, "need 1 <= max_batch <= max_streams");
    if (head_kind != VAPB_HEAD_VAP && head_kind != VAPB_HEAD_BC)
        return fail(nullptr, VAPB_EINVAL, "head_kind %d unknown", head_kind);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, VAPB_ECUDA, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(nullptr, VAPB_EINVAL, "device %d outside 0..%d", device, ndev - 1);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, VAPB_ECUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, VAPB_EUNSUPPORTED, "device %d is sm_%d%d; libvapb200 is built for sm_100a only", device,
                    prop.major, prop.minor);
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, VAPB_ECUDA, "cudaSetDevice failed");
    const int sm_count = prop.multiProcessorCount;

    vapb_ctx* c = new vapb_ctx();
    c->device = device; c->frame_hz = frame_hz; c->T = ctx_frames; c->max_streams = max_streams;
    c->max_batch = max_batch; c->head_kind = head_kind; c->sm_count = sm_count;
    c->S = 16000 / frame_hz + kPadSamples;
    const int ck[5] = {10, 8, 4, 4, 4}, cs[5] = {5, 4, 2, 2, 2}, cp[5] = {3, 2, 1, 1, 1};
    int len = c->S;
    for (int i = 0; i < 5; ++i) {
        len = (len + 2 * cp[i] - ck[i]) / cs[i] + 1;
        c->L[i] = len;
    }
    c->n_lstm = c->L[4] - 2;
    for (int i = 0; i < 4; ++i) c->halo[i] = cp[i + 1];   // halo rows of act[i] = padding of the conv that reads it
    c->halo[4] = 0;

#define FAIL_CREATE(code, ...)                       \
    do {                                             \
        int rc_ = fail(nullptr, code, __VA_ARGS__);  \
        vapb_destroy(c);                             \
        return rc_;                                  \
    } while (0)

    Loader ld;
    ld.c = c;
    if (!parse_blob(weights_blob, nbytes, ld.t, ld.err)) FAIL_CREATE(VAPB_EWEIGHTS, "%s", ld.err.c_str());
    const std::string G = "encoder.encoder.gEncoder.", A = "encoder.encoder.gAR.baseNet.";
    {   // conv0 weight [256][1][10] -> tap-major [10][256] so a warp reads it with coalesced float4 loads
        const HostTensor* h0 = ld.get(G + "conv0.weight", {256, 1, 10});
        std::vector<float> t(2560, 0.f);
        if (h0)
            for (int ch = 0; ch < 256; ++ch)
                for (int k = 0; k < 10; ++k) t[k * 256 + ch] = h0->data[ch * 10 + k];
        c->w0 = ld.upload(t);
    }
    c->b0 = ld.up(G + "conv0.bias", {256});
    c->cn0w = ld.up(G + "batchNorm0.weight", {1, 256, 1});
    c->cn0b = ld.up(G + "batchNorm0.bias", {1, 256, 1});
    for (int i = 0; i < 4; ++i) {
        ConvLayer& cv = c->conv[i];
        cv.k = ck[i + 1]; cv.s = cs[i + 1]; cv.p = cp[i + 1]; cv.Lin = c->L[i]; cv.Lout = c->L[i + 1];
        const std::string n = std::to_string(i + 1);
        cv.W = ld.upload(ld.conv_as_gemm(G + "conv" + n + ".weight", (uint32_t)cv.k));
        cv.b = ld.up(G + "conv" + n + ".bias", {256});
        cv.cnw = ld.up(G + "batchNorm" + n + ".weight", {1, 256, 1});
        cv.cnb = ld.up(G + "batchNorm" + n + ".bias", {1, 256, 1});
    }
    c->Wih = ld.up(A + "weight_ih_l0", {1024, 256});
    c->Whh = ld.up(A + "weight_hh_l0", {1024, 256});
    {
        const HostTensor* bi = ld.get(A + "bias_ih_l0", {1024});
        const HostTensor* bh = ld.get(A + "bias_hh_l0", {1024});
        std::vector<float> b(1024, 0.f);
        if (bi && bh)
            for (int i = 0; i < 1024; ++i) b[i] = bi->data[i] + bh->data[i];
        c->b_lstm = ld.upload(b);
    }
    {   // downsample Conv1d(256,256,k=n_lstm) as a GEMM over the n_lstm LSTM outputs (K = tap*256 + cin)
        const uint32_t kd = (uint32_t)c->n_lstm;
        c->Wds = ld.upload(ld.conv_as_gemm("encoder.downsample.1.weight", kd));
        c->bds = ld.up("encoder.downsample.1.bias", {256});
        c->ds_lnw = ld.up("encoder.downsample.2.ln.weight", {256});
        c->ds_lnb = ld.up("encoder.downsample.2.ln.bias", {256});
    }
    load_layer(ld, c->layers[0], "ar_channel.layers.0.", false);
    for (int i = 0; i < 3; ++i) load_layer(ld, c->layers[1 + i], "ar.layers." + std::to_string(i) + ".", true);
    c->Wa = ld.up("ar.combinator.h0_a.weight", {256, 256});
    c->Wb = ld.up("ar.combinator.h0_b.weight", {256, 256});
    c->comb_lnw = ld.up("ar.combinator.ln.weight", {256});
    c->comb_lnb = ld.up("ar.combinator.ln.bias", {256});
    if (head_kind == VAPB_HEAD_VAP) {
        c->n_out = 256;
        c->Wh = ld.up("vap_head.weight", {256, 256});
        c->bh = ld.up("vap_head.bias", {256});
        c->va_w = ld.up("va_classifier.weight", {1, 256});
        c->va_b = ld.up("va_classifier.bias", {1});
    } else {
        c->n_out = 3;
        c->Wh = ld.up("bc_head.weight", {3, 256});
        c->bh = ld.up("bc_head.bias", {3});
    }
    {
        const std::string p3 = "ar.layers.2.";
        c->t_WqT = ld.upT(p3 + "mha.query.weight", 256, 256);
        c->t_WprojT = ld.upT(p3 + "mha.proj.weight", 256, 256);
        c->t_WqcT = ld.upT(p3 + "mha_cross.query.weight", 256, 256);
        c->t_WprojcT = ld.upT(p3 + "mha_cross.proj.weight", 256, 256);
        c->t_W1T = ld.upT(p3 + "ffnetwork.0.weight", 768, 256);
        c->t_W2T = ld.upT(p3 + "ffnetwork.3.weight", 256, 768);
        c->t_WaT = ld.upT("ar.combinator.h0_a.weight", 256, 256);
        c->t_WbT = ld.upT("ar.combinator.h0_b.weight", 256, 256);
        c->t_WhT = head_kind == VAPB_HEAD_VAP ? ld.upT("vap_head.weight", 256, 256) : c->Wh;      // bc head: [3][256] as is
    }
    build_v2_weights(ld, c);
    if (!ld.ok) FAIL_CREATE(VAPB_EWEIGHTS, "%s", ld.err.c_str());

    // ---- state + workspaces
    const size_t MS = (size_t)max_streams, MB = (size_t)max_batch, NC = 2 * MB, R = NC * c->T;
    int rc = 0;
#define DA(ptr, n) if (!rc) rc = dalloc(c, &(ptr), (n))
    DA(c->hS, (MS + 1) * 2 * kD);      // + one scratch slot (index max_streams) for the bulk offline scorer
    DA(c->cS, (MS + 1) * 2 * kD);
    DA(c->bulk_ids, 1);
    DA(c->bulk_count, MB);
    DA(c->ring, MS * 2 * c->T * kD);
    DA(c->count, MS);
    DA(c->iobuf_dev, sizeof(IoPtrs) + MB * sizeof(int));
    DA(c->tvalid, MB);
    for (int i = 0; i < 5; ++i) DA(c->act[i], NC * (size_t)(c->L[i] + 2 * c->halo[i]) * kD);
    DA(c->hW, NC * kD);
    DA(c->cW, NC * kD);
    DA(c->Gx, NC * c->n_lstm * 4 * kD);
    DA(c->Gt, NC * 4 * kD);
    DA(c->Y, NC * c->n_lstm * kD);
    DA(c->dsout, NC * kD);
    DA(c->ebuf, NC * kD);
    DA(c->X, R * kD);
    DA(c->Z, R * kD);
    DA(c->QKV, R * 3 * kD);
    DA(c->O, R * kD);
    DA(c->Hd, R * kFF);
    DA(c->KVc, R * 2 * kD);
    DA(c->Qc, R * kD);
    c->part_stride = NC * (size_t)c->L[1] * kD;          // largest possible split-K user: conv1 output rows (option conv12_ks)
    DA(c->part, 4 * c->part_stride + 20 * NC * kD);      // conv3/conv4 use <= 4 slabs; the downsample up to 20 tiny ones
    DA(c->Xl, NC * kD);
    DA(c->Zl, NC * kD);
    DA(c->Ql, NC * kD);
    DA(c->Ol, NC * kD);
    DA(c->Hl, NC * kFF);
    DA(c->qkv_ring, MS * 2 * c->T * 3 * kD);
    DA(c->QKVn, NC * 3 * kD);
    DA(c->audio_stage, MB * 2 * c->S);
    DA(c->out_stage, MB * 6);
#undef DA
    if (rc) {
        g_create_error = c->err;
        vapb_destroy(c);
        return rc;
    }
    static_assert(sizeof(IoPtrs) == 16, "IoPtrs layout");
    c->io_dev = reinterpret_cast<IoPtrs*>(c->iobuf_dev);
    c->ids_dev = reinterpret_cast<int*>(c->iobuf_dev + sizeof(IoPtrs));
    if (cudaMallocHost(reinterpret_cast<void**>(&c->iobuf_pinned), sizeof(IoPtrs) + MB * sizeof(int)) != cudaSuccess)
        FAIL_CREATE(VAPB_ENOMEM, "cudaMallocHost failed");
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming);
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking) != cudaSuccess)
        FAIL_CREATE(VAPB_ECUDA, "cudaStreamCreate failed");
    cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);

    // ---- tcgen05 operands: bf16 hi/lo planes of every tensor-core GEMM weight (+ workspaces)
    {
        std::string terr;
        bool ok = true;
        for (int i = 0; i < 4 && ok; ++i) ok = tc_prepare_weight(c->conv[i].W, kD, c->conv[i].k * kD, c->conv[i].tc, c->allocs, terr);
        if (ok) ok = tc_prepare_weight(c->Wds, kD, c->n_lstm * kD, c->tc_ds, c->allocs, terr);
        if (ok) ok = tc_prepare_weight(c->Wih, 4 * kD, kD, c->tc_ih, c->allocs, terr);
        for (int l = 0; l < 4 && ok; ++l) {
            LayerWeights& lw = c->layers[l];
            ok = tc_prepare_weight(lw.sa.Wqkv, 3 * kD, kD, lw.sa.tc_qkv, c->allocs, terr) &&
                 tc_prepare_weight(lw.sa.Wqkv, kD, kD, lw.sa.tc_q, c->allocs, terr) &&
                 tc_prepare_weight(lw.sa.Wqkv + (size_t)kD * kD, 2 * kD, kD, lw.sa.tc_kv, c->allocs, terr) &&
                 tc_prepare_weight(lw.sa.Wproj, kD, kD, lw.sa.tc_proj, c->allocs, terr) &&
                 tc_prepare_weight(lw.W1, kFF, kD, lw.tc_w1, c->allocs, terr) &&
                 tc_prepare_weight(lw.W2, kD, kFF, lw.tc_w2, c->allocs, terr);
            if (ok && lw.cross)
                ok = tc_prepare_weight(lw.Wq_c, kD, kD, lw.tc_q_c, c->allocs, terr) &&
                     tc_prepare_weight(lw.Wkv_c, 2 * kD, kD, lw.tc_kv_c, c->allocs, terr) &&
                     tc_prepare_weight(lw.Wproj_c, kD, kD, lw.tc_proj_c, c->allocs, terr);
        }
        // largest A operand: conv1 input view (NC*L1 rows x 2048) or FFN hidden (R x 768)
        size_t max_a = std::max((size_t)NC * c->L[1] * c->conv[0].k * kD, R * (size_t)kFF);
        if (ok) ok = tc_prepare_workspace(c->tcws, max_a, c->allocs, terr);
        if (!ok) FAIL_CREATE(VAPB_ECUDA, "tcgen05 setup failed: %s", terr.c_str());
    }

    {
        const int scratch = max_streams;
        cudaMemcpy(c->bulk_ids, &scratch, sizeof(int), cudaMemcpyHostToDevice);
    }
    c->h_cnt.assign(MS, 0);
    c->h_qkv.assign(MS, 0);
    if (build_fused_ops(c) != 0) FAIL_CREATE(VAPB_ECUDA, "%s", c->err.c_str());
    if (build_fused2(c) != 0) FAIL_CREATE(VAPB_ECUDA, "%s", c->err.c_str());
    {   // weights the kernels behind the encoder read: bf16 planes of the transformer GEMMs, k-major tail weights
        std::vector<const void*> ptrs;
        std::vector<unsigned long long> bytes;
        auto add_tc = [&](const TcWeight& w) {
            if (!w.hi) return;
            ptrs.push_back(w.hi); bytes.push_back((unsigned long long)w.N * w.K * 2);
            ptrs.push_back(w.lo); bytes.push_back((unsigned long long)w.N * w.K * 2);
        };
        auto add_f = [&](const float* p, size_t n) { if (p) { ptrs.push_back(p); bytes.push_back((unsigned long long)n * 4); } };
        const bool v2 = c->f2ops != nullptr;
        for (int l = 0; l < 4; ++l) {
            const LayerWeights& lw = c->layers[l];
            if (v2) { add_tc(c->v2[l].g1); add_tc(c->v2[l].qc); add_tc(c->v2[l].w1); }
            else { add_tc(l < 3 ? lw.sa.tc_qkv : lw.sa.tc_kv); add_tc(lw.tc_kv_c); if (l < 3) { add_tc(lw.tc_q_c); add_tc(lw.tc_w1); } }
            if (l < 3) { add_tc(lw.sa.tc_proj); add_tc(lw.tc_proj_c); add_tc(lw.tc_w2); }
        }
        add_f(c->t_WqT, 65536); add_f(c->t_WprojT, 65536); add_f(c->t_WqcT, 65536); add_f(c->t_WprojcT, 65536);
        add_f(c->t_W1T, 196608); add_f(c->t_W2T, 196608); add_f(c->t_WaT, 65536); add_f(c->t_WbT, 65536);
        if (head_kind == VAPB_HEAD_VAP) add_f(c->t_WhT, 65536);
        c->pf_n = (int)ptrs.size();
        for (unsigned long long b : bytes) c->pf_lines += (b + 127) >> 7;
        void* d = nullptr;
        if (cudaMalloc(&d, ptrs.size() * sizeof(void*)) != cudaSuccess) FAIL_CREATE(VAPB_ENOMEM, "cudaMalloc failed");
        c->allocs.push_back(d);
        cudaMemcpy(d, ptrs.data(), ptrs.size() * sizeof(void*), cudaMemcpyHostToDevice);
        c->pf_ptrs = static_cast<const void**>(d);
        if (cudaMalloc(&d, bytes.size() * sizeof(unsigned long long)) != cudaSuccess) FAIL_CREATE(VAPB_ENOMEM, "cudaMalloc failed");
        c->allocs.push_back(d);
        cudaMemcpy(d, bytes.data(), bytes.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice);
        c->pf_bytes = static_cast<unsigned long long*>(d);
    }

    // taps (allocated lazily when keep_taps is switched on)
    if (cudaDeviceSynchronize() != cudaSuccess) FAIL_CREATE(VAPB_ECUDA, "device error during create: %s", cudaGetErrorString(cudaGetLastError()));
    *out = c;
    return VAPB_OK;
#undef FAIL_CREATE
}

int vapb_destroy(vapb_handle h) {
    if (!h) return VAPB_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto& g : h->graphs) cudaGraphExecDestroy(g.exec);
    for (void* p : h->allocs) cudaFree(p);
    if (h->iobuf_pinned) cudaFreeHost(h->iobuf_pinned);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    if (h->ev_out) cudaEventDestroy(h->ev_out);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    delete h;
    return VAPB_OK;
}

int vapb_chunk_samples(vapb_handle h) { return h ? h->S : VAPB_EINVAL; }

int vapb_reset_streams(vapb_handle h, const int* ids, int n) {
    if (!h) return VAPB_EINVAL;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    const size_t st = 2 * kD * sizeof(float), rg = (size_t)2 * h->T * kD * sizeof(float);
    if (!ids || n <= 0) {
        CK(h, cudaMemset(h->hS, 0, st * h->max_streams));
        CK(h, cudaMemset(h->cS, 0, st * h->max_streams));
        CK(h, cudaMemset(h->ring, 0, rg * h->max_streams));
        CK(h, cudaMemset(h->count, 0, sizeof(int) * h->max_streams));
        CK(h, cudaDeviceSynchronize());      // the memsets ran on the legacy stream: later steps may use any stream
        std::fill(h->h_cnt.begin(), h->h_cnt.end(), 0);
        std::fill(h->h_qkv.begin(), h->h_qkv.end(), 0);
        return VAPB_OK;
    }
    for (int i = 0; i < n; ++i) {
        const int id = ids[i];
        if (id < 0 || id >= h->max_streams) return fail(h, VAPB_EINVAL, "stream id %d outside 0..%d", id, h->max_streams - 1);
        CK(h, cudaMemset(h->hS + (size_t)id * 2 * kD, 0, st));
        CK(h, cudaMemset(h->cS + (size_t)id * 2 * kD, 0, st));
        CK(h, cudaMemset(h->ring + (size_t)id * 2 * h->T * kD, 0, rg));
        CK(h, cudaMemset(h->count + id, 0, sizeof(int)));
        h->h_cnt[id] = 0;
        h->h_qkv[id] = 0;
    }
    CK(h, cudaDeviceSynchronize());
    return VAPB_OK;
}

int vapb_step(vapb_handle h, const float* audio, const int* ids, int B, float* out, void* cuda_stream) {
    if (!h) return VAPB_EINVAL;
    if (!audio || !out) return fail(h, VAPB_EINVAL, "audio/out is NULL");
    int rc = check_ids(h, ids, B);
    if (rc) return rc;
    CK(h, cudaSetDevice(h->device));
    cudaStream_t caller = static_cast<cudaStream_t>(cuda_stream);
    cudaStream_t st = caller;
    const bool use_graph = h->opt_graph && !h->opt_keep_taps;
    const bool bridge = use_graph && (caller == nullptr || caller == cudaStreamLegacy || caller == cudaStreamPerThread);
    if (bridge) {
        st = h->own_stream;
        CK(h, cudaEventRecord(h->ev_in, caller));
        CK(h, cudaStreamWaitEvent(st, h->ev_in, 0));
    }
    rc = qkv_cache_pre_step(h, ids, B, st);
    if (rc) return rc;
    // {audio, out} + ids -> device in ONE copy.  The pinned staging buffer is reused, so wait for the previous copy first.
    if (h->last_B != 0) CK(h, cudaEventSynchronize(h->ev0));
    {
        IoPtrs io{audio, out};
        memcpy(h->iobuf_pinned, &io, sizeof io);
        memcpy(h->iobuf_pinned + sizeof io, ids, sizeof(int) * B);
        CK(h, cudaMemcpyAsync(h->iobuf_dev, h->iobuf_pinned, sizeof io + sizeof(int) * B, cudaMemcpyHostToDevice, st));
    }
    CK(h, cudaEventRecord(h->ev0, st));
    h->last_B = B;

    if (use_graph) {
        GraphEntry* ge = nullptr;
        for (auto& g : h->graphs)
            if (g.B == B) ge = &g;
        if (!ge) {
            Step s{h, st, B, nullptr, nullptr};
            s.io = h->io_dev;
            cudaGraph_t graph = nullptr;
            CK(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            enqueue_step(s);
            cudaError_t e = cudaStreamEndCapture(st, &graph);
            if (e != cudaSuccess || !graph) return fail(h, VAPB_ECUDA, "graph capture failed: %s", cudaGetErrorString(e));
            cudaGraphExec_t exec = nullptr;
            e = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) return fail(h, VAPB_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
            if (h->graphs.size() >= kMaxGraphs) {       // evict the least recently used batch size
                size_t lru = 0;
                for (size_t i = 1; i < h->graphs.size(); ++i)
                    if (h->graphs[i].last_used < h->graphs[lru].last_used) lru = i;
                cudaGraphExecDestroy(h->graphs[lru].exec);
                h->graphs.erase(h->graphs.begin() + (long)lru);
            }
            h->graphs.push_back(GraphEntry{B, exec, s.n, 0});
            ge = &h->graphs.back();
        }
        ge->last_used = ++h->tick;
        CK(h, cudaGraphLaunch(ge->exec, st));
        h->launches = ge->launches;
    } else {
        Step s{h, st, B, audio, out};
        enqueue_step(s);
        h->launches = s.n;
        CK(h, cudaGetLastError());
    }
    qkv_cache_post_step(h, ids, B);
    if (h->opt_timing) {
        CK(h, cudaEventRecord(h->ev1, st));
        h->timed = true;
    }
    if (bridge) {
        CK(h, cudaEventRecord(h->ev_out, st));
        CK(h, cudaStreamWaitEvent(caller, h->ev_out, 0));
    }
    return VAPB_OK;
}

int vapb_step_host(vapb_handle h, const float* audio, const int* ids, int B, float* out, void* cuda_stream) {
    if (!h) return VAPB_EINVAL;
    if (B <= 0 || B > h->max_batch) return fail(h, VAPB_EINVAL, "B=%d outside 1..%d", B, h->max_batch);
    if (!audio || !out) return fail(h, VAPB_EINVAL, "audio/out is NULL");
    CK(h, cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    CK(h, cudaMemcpyAsync(h->audio_stage, audio, sizeof(float) * (size_t)B * 2 * h->S, cudaMemcpyHostToDevice, st));
    int rc = vapb_step(h, h->audio_stage, ids, B, h->out_stage, cuda_stream);
    if (rc) return rc;
    CK(h, cudaMemcpyAsync(out, h->out_stage, sizeof(float) * (size_t)B * 6, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    return VAPB_OK;
}

int vapb_score_offline(vapb_handle h, const float* audio, long long n_samples, float* out, long long max_frames, long long* n_frames,
                       void* cuda_stream) {
    if (!h) return VAPB_EINVAL;
    if (!audio || !out || !n_frames) return fail(h, VAPB_EINVAL, "null argument");
    CK(h, cudaSetDevice(h->device));
    vapb_ctx* c = h;
    const int S = c->S, shift = S - kPadSamples, T = c->T, MB = c->max_batch;
    const long long N = n_samples >= S ? (n_samples - S) / shift + 1 : 0;
    *n_frames = N;
    if (N == 0) return VAPB_OK;
    if (N > max_frames) return fail(h, VAPB_EINVAL, "out holds %lld frames, the audio has %lld", max_frames, N);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if (st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread) st = c->own_stream;
    CK(h, cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream)));
    float* E = nullptr;                    // [2][N][256] embeddings of the whole file
    if (cudaMalloc(&E, (size_t)2 * N * kD * sizeof(float)) != cudaSuccess) return fail(h, VAPB_ENOMEM, "cudaMalloc(embeddings) failed");
    vapb::g_use_pdl = false;
    vapb::g_attn_rk = c->opt_attn_rk != 0;
    // ---- pass 1: encoder.  The conv stack treats every chunk in isolation (zero padding per chunk, encoder.py:58-80), so
    // max_batch consecutive chunks of BOTH channels go through it as one batch; only the LSTM is sequential in time: one
    // cluster walks n_lstm * Nc steps per channel with the state of a scratch slot that starts at zero (a fresh VAPRealTime).
    CK(h, cudaMemsetAsync(c->hS + (size_t)c->max_streams * 2 * kD, 0, 2 * kD * sizeof(float), st));
    CK(h, cudaMemsetAsync(c->cS + (size_t)c->max_streams * 2 * kD, 0, 2 * kD * sizeof(float), st));
    for (long long b0 = 0; b0 < N; b0 += MB) {
        const int Nc = (int)std::min<long long>(MB, N - b0);
        launch_make_chunks(audio, n_samples, shift, S, b0, Nc, c->audio_stage, st);
        Step s{c, st, Nc, c->audio_stage, nullptr};
        encoder_convs(s, 2 * Nc);                                  // chunk rows are channel-major: row = ch * Nc + b
        launch_lstm_recurrent(c->Gx, c->Whh, c->hS, c->cS, c->bulk_ids, c->Y, 2, c->n_lstm * Nc, st);
        const int Kd = c->n_lstm * kD;
        const int ks = (c->opt_gemm == 1 && c->opt_splitk) ? (Kd / 64) / 2 : 1;
        const long long dstride = (long long)2 * Nc * kD;
        if (ks > 1) {
            gemm(s, "gemm_downsample", c->Y, plain_map(Kd), c->Wds, &c->tc_ds, c->bds, nullptr, plain_map(kD), c->part, plain_map(kD), 2 * Nc, kD, Kd, 0,
                 nullptr, nullptr, ks, dstride, c->opt_conv4p);
            launch_ln_gelu_ring(c->part, Nc, c->ds_lnw, c->ds_lnb, nullptr, nullptr, nullptr, T, c->ebuf, st, ks, dstride);
        } else {
            gemm(s, "gemm_downsample", c->Y, plain_map(Kd), c->Wds, &c->tc_ds, c->bds, nullptr, plain_map(kD), c->dsout, plain_map(kD), 2 * Nc, kD, Kd, 0);
            launch_ln_gelu_ring(c->dsout, Nc, c->ds_lnw, c->ds_lnb, nullptr, nullptr, nullptr, T, c->ebuf, st);
        }
        for (int ch = 0; ch < 2; ++ch)
            CK(h, cudaMemcpyAsync(E + ((size_t)ch * N + b0) * kD, c->ebuf + (size_t)ch * Nc * kD, (size_t)Nc * kD * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    // ---- pass 2: the window of frame f is e[f-T+1 .. f]; windows are independent, so max_batch of them form one batch of
    // the batched transformer kernels (last layer pruned to the newest frame, heads on the newest frame)
    const bool prune = c->opt_prune && !c->opt_keep_taps;
    for (long long f0 = 0; f0 < N; f0 += MB) {
        const int Bw = (int)std::min<long long>(MB, N - f0);
        launch_gather_windows(E, N, f0, Bw, T, c->X, c->tvalid, c->ids_dev, st);
        Step s{c, st, Bw, nullptr, out + (size_t)f0 * 6};
        transformer_batched(s, prune, false);
        tail_or_head(s, prune, c->bulk_count);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(E);
    h->last_B = 0;                          // ids_dev / tvalid were used as scratch
    if (e != cudaSuccess) return fail(h, VAPB_ECUDA, "offline scoring failed: %s", cudaGetErrorString(e));
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, VAPB_ECUDA, "offline scoring failed: %s", cudaGetErrorString(e));
    return VAPB_OK;
}

size_t vapb_state_floats(vapb_handle h) { return h ? (size_t)2 + 4 * kD + (size_t)2 * h->T * kD : 0; }

int vapb_export_state(vapb_handle h, int id, float* state) {
    if (!h || !state) return VAPB_EINVAL;
    if (id < 0 || id >= h->max_streams) return fail(h, VAPB_EINVAL, "stream id %d outside 0..%d", id, h->max_streams - 1);
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    int cnt = 0;
    CK(h, cudaMemcpy(&cnt, h->count + id, sizeof(int), cudaMemcpyDeviceToHost));
    const int T = h->T, t = std::min(cnt, T);
    state[0] = (float)cnt;
    state[1] = (float)t;
    CK(h, cudaMemcpy(state + 2, h->hS + (size_t)id * 2 * kD, 2 * kD * sizeof(float), cudaMemcpyDeviceToHost));
    CK(h, cudaMemcpy(state + 2 + 2 * kD, h->cS + (size_t)id * 2 * kD, 2 * kD * sizeof(float), cudaMemcpyDeviceToHost));
    std::vector<float> raw((size_t)2 * T * kD);
    CK(h, cudaMemcpy(raw.data(), h->ring + (size_t)id * 2 * T * kD, raw.size() * sizeof(float), cudaMemcpyDeviceToHost));
    float* r = state + 2 + 4 * kD;
    memset(r, 0, raw.size() * sizeof(float));
    for (int ch = 0; ch < 2; ++ch)
        for (int j = 0; j < t; ++j) {
            const int slot = (cnt - t + j) % T;
            memcpy(r + ((size_t)ch * T + j) * kD, raw.data() + ((size_t)ch * T + slot) * kD, kD * sizeof(float));
        }
    return VAPB_OK;
}

int vapb_import_state(vapb_handle h, int id, const float* state) {
    if (!h || !state) return VAPB_EINVAL;
    if (id < 0 || id >= h->max_streams) return fail(h, VAPB_EINVAL, "stream id %d outside 0..%d", id, h->max_streams - 1);
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    const int T = h->T;
    const int cnt = (int)state[0], t = (int)state[1];
    if (cnt < 0 || t != std::min(cnt, T)) return fail(h, VAPB_EINVAL, "inconsistent state record (count %d, t %d, T %d)", cnt, t, T);
    CK(h, cudaMemcpy(h->count + id, &cnt, sizeof(int), cudaMemcpyHostToDevice));
    CK(h, cudaMemcpy(h->hS + (size_t)id * 2 * kD, state + 2, 2 * kD * sizeof(float), cudaMemcpyHostToDevice));
    CK(h, cudaMemcpy(h->cS + (size_t)id * 2 * kD, state + 2 + 2 * kD, 2 * kD * sizeof(float), cudaMemcpyHostToDevice));
    std::vector<float> raw((size_t)2 * T * kD, 0.f);
    const float* r = state + 2 + 4 * kD;
    for (int ch = 0; ch < 2; ++ch)
        for (int j = 0; j < t; ++j) {
            const int slot = (cnt - t + j) % T;
            memcpy(raw.data() + ((size_t)ch * T + slot) * kD, r + ((size_t)ch * T + j) * kD, kD * sizeof(float));
        }
    CK(h, cudaMemcpy(h->ring + (size_t)id * 2 * T * kD, raw.data(), raw.size() * sizeof(float), cudaMemcpyHostToDevice));
    h->h_cnt[id] = cnt;
    h->h_qkv[id] = -1;           // cached layer-0 projections no longer match the ring: rebuilt before the next batched step
    return VAPB_OK;
}

int vapb_set_option(vapb_handle h, const char* key, int value) {
    if (!h || !key) return VAPB_EINVAL;
    const std::string k(key);
    if (k == "graph") h->opt_graph = value ? 1 : 0;
    else if (k == "gemm") {
        if (value != 0 && value != 1) return fail(h, VAPB_EINVAL, "gemm must be 0 or 1");
        if (value != h->opt_gemm) {
            cudaSetDevice(h->device);
            cudaDeviceSynchronize();
            for (auto& g : h->graphs) cudaGraphExecDestroy(g.exec);
            h->graphs.clear();
        }
        h->opt_gemm = value;
    } else if (k == "timing") h->opt_timing = value ? 1 : 0;
    else if (k == "lstm_fused" || k == "tile_n" || k == "fuse_ln" || k == "k256" || k == "pdl" || k == "prune" || k == "attn_rk" || k == "fork" || k == "splitk" || k == "conv4p" || k == "cluster2" || k == "fused" || k == "fused_dbg" || k == "fused_v" || k == "tail" || k == "lstm_x_tc" || k == "qkv_cache" || k == "conv12_ks" || k == "prefetch") {
        if (k == "tile_n" && value != 0 && value != 64 && value != 128 && value != 256) return fail(h, VAPB_EINVAL, "tile_n must be 0, 64, 128 or 256");
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        for (auto& g : h->graphs) cudaGraphExecDestroy(g.exec);
        h->graphs.clear();
        if (k == "lstm_fused") h->opt_lstm_fused = value ? 1 : 0;
        else if (k == "fuse_ln") h->opt_fuse_ln = value ? 1 : 0;
        else if (k == "k256") h->tcws.use_k256 = value ? 1 : 0;
        else if (k == "pdl") h->opt_pdl = value ? 1 : 0;
        else if (k == "prune") h->opt_prune = value ? 1 : 0;
        else if (k == "attn_rk") h->opt_attn_rk = value ? 1 : 0;
        else if (k == "fork") h->opt_fork = value ? 1 : 0;
        else if (k == "splitk") h->opt_splitk = value ? 1 : 0;
        else if (k == "conv4p") h->opt_conv4p = value & 3;
        else if (k == "cluster2") h->tcws.cluster2 = value ? 1 : 0;
        else if (k == "fused") h->opt_fused = value;      // 0 off, 1 auto (one wave of clusters), 2 always
        else if (k == "fused_dbg") h->opt_fused_dbg = value;
        else if (k == "fused_v") h->opt_fused_v = value == 1 ? 1 : 2;
        else if (k == "tail") h->opt_tail = value ? 1 : 0;
        else if (k == "lstm_x_tc") h->opt_lstm_x_tc = value ? 1 : 0;
        else if (k == "qkv_cache") h->opt_qkv_cache = value ? 1 : 0;
        else if (k == "conv12_ks") h->opt_conv12_ks = value & 0xff;
        else if (k == "prefetch") h->opt_prefetch = value ? 1 : 0;
        else { h->opt_tile_n = value; h->tcws.force_bn = value; }
    } else if (k == "keep_taps") {
        h->opt_keep_taps = value ? 1 : 0;
        if (value && !h->tap_chan) {
            CK(h, cudaSetDevice(h->device));
            const size_t RX = (size_t)2 * h->max_batch * h->T * kD;
            int rc = dalloc(h, &h->tap_xin, RX);
            if (!rc) rc = dalloc(h, &h->tap_chan, RX);
            for (int i = 0; i < 3 && !rc; ++i) rc = dalloc(h, &h->tap_cross[i], RX);
            if (!rc) rc = dalloc(h, &h->tap_comb, (size_t)h->max_batch * kD);
            if (!rc) rc = dalloc(h, &h->tap_logits, (size_t)h->max_batch * kD);
            if (rc) return rc;
        }
    } else return fail(h, VAPB_EINVAL, "unknown option %s", key);
    return VAPB_OK;
}

int vapb_get_option(vapb_handle h, const char* key, int* value) {
    if (!h || !key || !value) return VAPB_EINVAL;
    const std::string k(key);
    if (k == "graph") *value = h->opt_graph;
    else if (k == "gemm") *value = h->opt_gemm;
    else if (k == "timing") *value = h->opt_timing;
    else if (k == "lstm_fused") *value = h->opt_lstm_fused;
    else if (k == "tile_n") *value = h->opt_tile_n;
    else if (k == "fuse_ln") *value = h->opt_fuse_ln;
    else if (k == "k256") *value = h->tcws.use_k256;
    else if (k == "pdl") *value = h->opt_pdl;
    else if (k == "prune") *value = h->opt_prune;
    else if (k == "attn_rk") *value = h->opt_attn_rk;
    else if (k == "fork") *value = h->opt_fork;
    else if (k == "splitk") *value = h->opt_splitk;
    else if (k == "conv4p") *value = h->opt_conv4p;
    else if (k == "cluster2") *value = h->tcws.cluster2;
    else if (k == "fused") *value = h->opt_fused;
    else if (k == "fused_dbg") *value = h->opt_fused_dbg;
    else if (k == "tail") *value = h->opt_tail;
    else if (k == "lstm_x_tc") *value = h->opt_lstm_x_tc;
    else if (k == "qkv_cache") *value = h->opt_qkv_cache;
    else if (k == "conv12_ks") *value = h->opt_conv12_ks;
    else if (k == "prefetch") *value = h->opt_prefetch;
    else if (k == "fused_v") *value = (h->opt_fused_v == 2 && h->f2ops) ? 2 : 1;
    else if (k == "keep_taps") *value = h->opt_keep_taps;
    else return fail(h, VAPB_EINVAL, "unknown option %s", key);
    return VAPB_OK;
}

int vapb_debug_tensor(vapb_handle h, const char* name, float* host_out, size_t cap, size_t* n_out) {
    if (!h || !name || !n_out) return VAPB_EINVAL;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    const int B = h->last_B, NC = 2 * B;
    if (B == 0) return fail(h, VAPB_EINVAL, "no step has run yet");
    const std::string k(name);
    const size_t RX = (size_t)NC * h->T * kD;
    const float* src = nullptr;
    size_t n = 0;
    std::vector<float> tmp;
    if (k.size() == 5 && k.compare(0, 4, "conv") == 0 && k[4] >= '0' && k[4] <= '4') {
        // strip the halo rows: [NC][L][256]
        const int i = k[4] - '0';
        const size_t rows = (size_t)h->L[i] + 2 * h->halo[i];
        std::vector<float> raw((size_t)NC * rows * kD);
        CK(h, cudaMemcpy(raw.data(), h->act[i], raw.size() * sizeof(float), cudaMemcpyDeviceToHost));
        tmp.resize((size_t)NC * h->L[i] * kD);
        for (int c = 0; c < NC; ++c)
            memcpy(tmp.data() + (size_t)c * h->L[i] * kD, raw.data() + ((size_t)c * rows + h->halo[i]) * kD,
                   (size_t)h->L[i] * kD * sizeof(float));
        n = tmp.size();
    } else if (k == "fused_clocks") {
        // clock64 deltas (cycles) per op of the stream kernel, cluster 0 / CTA 0 (option fused_dbg)
        std::vector<long long> clk(64);
        CK(h, cudaMemcpy(clk.data(), h->fused_clk, 64 * sizeof(long long), cudaMemcpyDeviceToHost));
        const int nops = (h->opt_fused_v == 2 && h->f2ops) ? h->n_f2ops : h->n_fops;
        tmp.resize((size_t)nops);
        for (int i = 0; i < nops; ++i) tmp[i] = (float)(clk[i + 1] - clk[i]);
        for (int i = 1; i < 13; ++i) tmp.push_back(clk[40 + i] ? (float)(clk[40 + i] - clk[40]) : 0.f);   // fine stamps of op fused_dbg - 1
        n = tmp.size();
    } else if (k == "lstm_out") { src = h->Y; n = (size_t)NC * h->n_lstm * kD; }
    else if (k == "e") { src = h->ebuf; n = (size_t)NC * kD; }
    else if (k == "x_out") { src = h->X; n = RX; }
    else if (h->opt_keep_taps && k == "x_in") { src = h->tap_xin; n = RX; }
    else if (h->opt_keep_taps && k == "chan_out") { src = h->tap_chan; n = RX; }
    else if (h->opt_keep_taps && k.size() == 10 && k.compare(0, 5, "cross") == 0 && k.compare(6, 4, "_out") == 0 && k[5] >= '0' && k[5] <= '2') {
        src = h->tap_cross[k[5] - '0']; n = RX;
    } else if (h->opt_keep_taps && k == "comb") { src = h->tap_comb; n = (size_t)B * kD; }
    else if (h->opt_keep_taps && k == "logits") { src = h->tap_logits; n = (size_t)B * kD; }
    else return fail(h, VAPB_EINVAL, "unknown tap %s (or keep_taps is off)", name);
    *n_out = n;
    if (!host_out) return VAPB_OK;
    if (cap < n) return fail(h, VAPB_EINVAL, "tap %s needs %zu floats, buffer has %zu", name, n, cap);
    if (src) CK(h, cudaMemcpy(host_out, src, n * sizeof(float), cudaMemcpyDeviceToHost));
    else memcpy(host_out, tmp.data(), n * sizeof(float));
    return VAPB_OK;
}

int vapb_last_launch_count(vapb_handle h) { return h ? h->launches : VAPB_EINVAL; }

int vapb_last_step_ms(vapb_handle h, float* ms) {
    if (!h || !ms) return VAPB_EINVAL;
    if (!h->timed) return fail(h, VAPB_EINVAL, "option timing is off or no step has run");
    CK(h, cudaEventSynchronize(h->ev1));
    CK(h, cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return VAPB_OK;
}

int vapb_profile_step(vapb_handle h, const float* audio, const int* ids, int B, float* out, void* cuda_stream,
                      char* report, size_t cap) {
    if (!h) return VAPB_EINVAL;
    if (!audio || !out || !report || cap == 0) return fail(h, VAPB_EINVAL, "null argument");
    int rc = check_ids(h, ids, B);
    if (rc) return rc;
    CK(h, cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    CK(h, cudaStreamSynchronize(st));
    rc = qkv_cache_pre_step(h, ids, B, st);
    if (rc) return rc;
    memcpy(h->iobuf_pinned + sizeof(IoPtrs), ids, sizeof(int) * B);
    CK(h, cudaMemcpyAsync(h->ids_dev, h->iobuf_pinned + sizeof(IoPtrs), sizeof(int) * B, cudaMemcpyHostToDevice, st));
    h->last_B = B;
    std::vector<std::pair<const char*, cudaEvent_t>> ev;
    Step s{h, st, B, audio, out};
    s.prof = &ev;
    cudaEvent_t e0;
    cudaEventCreate(&e0);
    cudaEventRecord(e0, st);
    CK(h, cudaEventRecord(h->ev0, st));
    enqueue_step(s);
    h->launches = s.n;
    qkv_cache_post_step(h, ids, B);
    cudaError_t e = cudaStreamSynchronize(st);
    std::vector<std::pair<std::string, std::pair<float, int>>> agg;
    cudaEvent_t prev = e0;
    float total = 0.f;
    for (auto& pr : ev) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, prev, pr.second);
        total += ms;
        bool found = false;
        for (auto& a : agg)
            if (a.first == pr.first) { a.second.first += ms; a.second.second += 1; found = true; }
        if (!found) agg.push_back({pr.first, {ms, 1}});
        prev = pr.second;
    }
    cudaEventDestroy(e0);
    for (auto& pr : ev) cudaEventDestroy(pr.second);
    if (e != cudaSuccess) return fail(h, VAPB_ECUDA, "profile step failed: %s", cudaGetErrorString(e));
    std::string rep = "tag,launches,ms\n";
    char line[160];
    for (auto& a : agg) {
        snprintf(line, sizeof line, "%s,%d,%.6f\n", a.first.c_str(), a.second.second, a.second.first);
        rep += line;
    }
    snprintf(line, sizeof line, "total,%d,%.6f\n", (int)ev.size(), total);
    rep += line;
    snprintf(report, cap, "%s", rep.c_str());
    return VAPB_OK;
}

int vapb_selftest_gemm(int device, int variant, double* max_rel_err) {
    std::string err;
    int rc = tc_selftest(device, variant, max_rel_err, err);
    g_create_error = err;      // the report (also on success) is readable through vapb_last_error(NULL)
    return rc;
}

}  // extern "C"
