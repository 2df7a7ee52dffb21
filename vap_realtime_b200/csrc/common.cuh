// Shared declarations of libvapb200 (B200 / sm_100a VAP streaming step).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace vapb {

constexpr int kD = 256;        // model width            (rvap/vap_main/vap_main.py:50)
constexpr int kFF = 768;       // FFN width, dff_k = 3   (vap_main.py:110)
constexpr int kHeads = 4;      // heads                  (vap_main.py:53)
constexpr int kHeadDim = 64;
constexpr int kPadSamples = 320;   // frame_contxt_padding (vap_main.py:224)
constexpr int kMaxT = 128;     // largest supported window (frames)
constexpr float kEps = 1e-5f;

// Affine row addressing used for activations that are stored per audio chunk
// with zero "halo" rows (the conv padding materialised once at allocation):
//   element offset of logical row r = offset + (r / rpc) * chunk_stride + (r % rpc) * row_stride
// A plain [M, ld] matrix is {rpc = 1<<30, chunk_stride = 0, row_stride = ld, offset = 0}.
struct RowMap {
    int rpc;
    long long chunk_stride;
    long long row_stride;
    long long offset;
};

__host__ __device__ inline long long rowmap_off(const RowMap& m, int r) {
    if (m.rpc >= (1 << 30)) return m.offset + (long long)r * m.row_stride;      // plain matrix: no division
    const int q = r / m.rpc;
    return m.offset + (long long)q * m.chunk_stride + (long long)(r - q * m.rpc) * m.row_stride;
}

inline RowMap plain_map(long long ld, long long offset = 0) {
    RowMap m;
    m.rpc = 1 << 30;
    m.chunk_stride = 0;
    m.row_stride = ld;
    m.offset = offset;
    return m;
}

// Caller-owned buffers of one step, read by the kernels THROUGH device memory when the step runs as a CUDA graph:
// the graph is captured once per batch size and vapb_step only rewrites these two pointers (same small H2D copy
// that carries the stream ids), so callers may pass any audio / out buffer without a re-capture.
struct IoPtrs {
    const float* audio;   // [B][2][S]
    float* out;           // [B][6]
};

struct GemmArgs {
    const float* A;
    RowMap amap;          // row m of A (K contiguous floats)
    const float* W;       // [N][K] row major (K contiguous): C = A * W^T
    const float* bias;    // [N] or nullptr
    const float* R;       // residual added after the activation, or nullptr
    RowMap rmap;
    float* C;
    RowMap cmap;
    int M, N, K;
    int act;              // 0 = none, 1 = exact-erf GELU
    // Optional LayerNorm(256) prologue applied to the rows of A before the product (tcgen05 path,
    // K == 256 only): A_norm = (A - mean) * rsqrt(var + 1e-5) * ln_w + ln_b  (biased variance).
    const float* ln_w = nullptr;
    const float* ln_b = nullptr;
    long long* dbg = nullptr;     // optional: clock64 stamps per CTA (self-test / tuning only)
    // split-K (tcgen05 generic kernel): grid.z = ksplit CTAs each take K/ksplit; partial z is written to
    // C + z * csplit_stride (no activation / residual; bias only in partial 0); the consumer sums them.
    int ksplit = 1;
    long long csplit_stride = 0;
    // Extra-precision modes of the generic tcgen05 kernel, used for the conv stack whose rounding is
    // amplified by the ChannelNorms and dominates the end-to-end error (precision):
    //   bit 0: keep the small products (hi*lo, lo*hi[, lo*lo]) in a second TMEM accumulator and add it to
    //          the hi*hi accumulator in fp32 in the epilogue (no extra MMA);
    //   bit 1: also issue the lo*lo product (4 MMAs per k-step instead of 3).
    int four_products = 0;
};

// ---- programmatic dependent launch (PDL) ---------------------------------------------------
// Every kernel of the step calls pdl_trigger() first (lets the NEXT kernel of the stream start its
// prologue: barrier init, TMEM allocation, tensor-map prefetch, weight TMA) and pdl_wait() before it
// touches anything a previous kernel produced.  Both are no-ops for a normal launch.
#ifdef __CUDACC__
#ifdef VAPB_ENABLE_PDL
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#else
// PDL measured slower than plain graph edges on this driver (profiles/r01_*_option_ablation.log) and
// griddepcontrol.wait costs ~1k cycles even on a normal launch, so it is compiled out by default.
__device__ __forceinline__ void pdl_trigger() {}
__device__ __forceinline__ void pdl_wait() {}
#endif

// cudaFuncSetAttribute is per device: true the first time a given call site runs on the current device
struct OncePerDevice {
    bool done[64] = {};
    bool first() {
        int d = 0;
        cudaGetDevice(&d);
        d &= 63;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};

extern thread_local bool g_use_pdl;      // set by the step driver before it enqueues kernels
extern thread_local bool g_attn_rk;

// (A uniform max-shared-memory carve-out hint for every kernel was tried and measured ~5 % slower:
//  the row-wise kernels lose their L1.  profiles/r01_j_option_ablation.log)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
#ifdef VAPB_ENABLE_PDL
    cfg.numAttrs = g_use_pdl ? 1 : 0;
#else
    cfg.numAttrs = 0;
#endif
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// launch with a thread-block cluster of `cluster_x` CTAs along x
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                                    Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cluster_x;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---- launchers implemented in kernels_simt.cu -------------------------------------------
void launch_conv0(const float* audio, const IoPtrs* io, int n_chunks, int S, int L0, const float* w, const float* b,
                  const float* cnw, const float* cnb, float* out, RowMap omap, cudaStream_t st);      // io != null: audio = io->audio
void launch_sgemm(const GemmArgs& g, cudaStream_t st);
void launch_cn_relu(float* X, RowMap map, int M, const float* w, const float* b, cudaStream_t st,
                    const float* partials = nullptr, int nsplit = 0, long long split_stride = 0);
void launch_layernorm(const float* X, RowMap xmap, float* Y, RowMap ymap, int M, const float* w,
                      const float* b, int gelu, cudaStream_t st);
void launch_gather_state(const float* hS, const float* cS, const int* ids, float* hW, float* cW, int B,
                         cudaStream_t st);
void launch_scatter_state(float* hS, float* cS, const int* ids, const float* hW, const float* cW, int B,
                          cudaStream_t st);
void launch_lstm_cell(const float* G, float* hW, float* cW, float* Y, int n_rows, int n_steps, int step,
                      cudaStream_t st);
void launch_lstm_recurrent(const float* Gx, const float* Whh, float* hS, float* cS, const int* ids, float* Y, int NC,
                           int n_steps, cudaStream_t st);
void launch_ln_gelu_ring(const float* X, int B, const float* w, const float* b, float* ring,
                         const int* count, const int* ids, int T, float* e_out, cudaStream_t st,
                         int nsplit = 1, long long split_stride = 0);
void launch_gather_ring(const float* ring, const int* count, const int* ids, float* X, int* tvalid, int B,
                        int T, cudaStream_t st);
struct AttnArgs {
    const float* Q; int ldq;       // row r, head h at Q + r*ldq + h*64
    const float* K; int ldk;
    const float* V; int ldv;
    float* O; int ldo;
    const int* tvalid;             // [B] valid rows per stream
    const float* slopes;           // [4] ALiBi m_h
    int n_seq;                     // 2B sequences of T rows each
    int T;
    int sibling;                   // 1: K/V rows come from the other channel of the same stream
    // Layer-0 Q/K/V cache: when ring_ids != null, Q / K / V point at the per-stream ring [max_streams][2][T][ld] and the row of
    // logical position j of sequence n is ring slot (count + 1 - t + j) % T of stream ring_ids[n / 2] (O stays in batch order)
    const int* ring_ids = nullptr;
    const int* ring_count = nullptr;
};
void launch_attention(const AttnArgs& a, cudaStream_t st);
// rows [n][768] of freshly projected Q|K|V of the newest frame -> ring slot count % T of their streams
void launch_qkv_append(const float* qkv_new, float* qkv_ring, const int* count, const int* ids, int B, int T, cudaStream_t st);
// last-row-only variants used when the final cross layer is pruned to the newest frame
void launch_gather_last(const float* X, const int* tvalid, float* Xl, int n_seq, int T, cudaStream_t st);
void launch_attention_last(const AttnArgs& a, cudaStream_t st);   // Q, O: [n_seq][256] compact; K, V: full rows
void launch_vad(const float* X, const int* tvalid, const float* w, const float* b, float* out, const IoPtrs* io, int B, int T,
                cudaStream_t st);
struct HeadArgs {
    const float* X;            // [2B*T][256] final cross-layer output
    const int* tvalid;
    const float* Wa; const float* Wb; const float* lnw; const float* lnb;
    const float* Wh; const float* bh; int n_out;   // 256 (vap) or 3 (bc)
    float* out;                // [B][6]
    const IoPtrs* io;          // != null: out = io->out
    float* comb_tap;           // [B][256] or nullptr
    float* logits_tap;         // [B][256] or nullptr
    int* count; const int* ids;
    int B, T, head_kind;
    int compact;               // 1: X is [2B][256] (one row per sequence: the newest frame)
};
void launch_head(const HeadArgs& a, cudaStream_t st);

// Newest-frame tail of the pruned last cross layer + Combinator + head in ONE kernel (kernels_simt.cu: k_tail).
// All weights are k-major transposes ([K][N]) so that a warp reads 128 contiguous bytes of one weight row slice.
struct TailArgs {
    const float* Xl;                 // [2B][256] newest frame of every sequence = input rows of the last layer
    const float* KVs;                // self-attention keys | values of the window: row (n * T + j), 512 floats
    const float* KVc;                // cross-attention keys | values (projected from the raw sibling rows), same layout
    const int* tvalid;               // [B]
    const float *WqT, *WprojT, *WqcT, *WprojcT;      // [256][256]
    const float *W1T;                // [256][768]
    const float *W2T;                // [768][256]
    const float *WaT, *WbT;          // combinator, [256][256]
    const float *WhT;                // vap head [256][256] (k-major); bc head: [3][256] row-major
    const float *ln_sa_w, *ln_sa_b, *ln_src_w, *ln_src_b, *ln_ff_w, *ln_ff_b, *comb_lnw, *comb_lnb, *bh;
    const float *slopes_s, *slopes_c;
    int n_out, head_kind, B, T;
    float* out;
    const IoPtrs* io;
    int* count;
    const int* ids;
};
void launch_tail(const TailArgs& a, cudaStream_t st);

// pulls `n` buffers (device arrays of pointers / byte counts) into L2, one prefetch per 128-byte line
void launch_l2_prefetch(const void* const* ptrs, const unsigned long long* bytes, int n, unsigned long long total_lines, cudaStream_t st);

// ---- bulk offline scoring (vap_offline.py:51-73 over a whole file) ----
// chunk rows, channel-major: dst[(ch * n_chunks + b)][0..S) = audio[ch][shift * (first + b) .. + S)
void launch_make_chunks(const float* audio, long long n_samples, int shift, int S, long long first, int n_chunks, float* dst, cudaStream_t st);
// X[(2b + ch) * T + j] = E[ch][f - t + 1 + j] for j < t = min(f + 1, T), f = first + b; zero rows above; tvalid[b] = t
void launch_gather_windows(const float* E, long long n_frames, long long first, int B, int T, float* X, int* tvalid, int* ids_out, cudaStream_t st);

}  // namespace vapb
