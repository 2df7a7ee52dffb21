// Per-stream persistent transformer kernel, second generation ("stream kernel v2"), T <= 64.
//
// A cluster of FOUR CTAs owns two stereo streams for the whole stack: CTA (stream slot s, column half r).  The M = 128 tile
// of a stream is both of its channels (sequence c in tile rows 64c .. 64c + T - 1); CTA r computes output columns
// 256 j + 128 r .. + 127 of every GEMM, i.e. heads 2r, 2r + 1 of every projection.  Every operand reaches the tensor core
// the standard way -- TMA -> shared memory -> tcgen05.mma -- and every conversion happens ONCE, in the epilogue of the op
// that produces a tensor:
//
//   * activations live in L2 as bf16 hi / lo planes in tile-row layout ([slot * 128 rows][cols]); the consuming op fetches
//     128 x 64 tiles with TMA (128 B swizzle), the two CTAs of a stream each multicast half of every A tile, the two CTAs of
//     a column half each multicast half of every W tile.  All 512 TMEM columns are accumulators (4 x 128): pairs of
//     subtiles are issued interleaved, a single subtile as one chain of N = 128 MMAs;
//   * the epilogue of the producing op writes the planes: tensor memory hands a thread one row, the rows of a 32-column
//     block pass through the warp's staging tile (two SWIZZLE_64B half tiles: hi, lo) and leave as TMA stores whose maps
//     are [sequence][position < T][cols], so the padding rows of a tile are never written;
//   * ONE 1 280-column scratch row per tile row carries every intermediate of a layer ([Q | K | V | K_c | V_c], the
//     attention output over the Q columns, the cross attention's Q and output, the FFN hidden rows): the working set of
//     64 streams fits L2;
//   * LayerNorm moved from the prologue to the epilogue: the GEMM runs on the RAW rows with weights pre-scaled by the
//     LayerNorm gain, y_n = rstd * (x (W o g)^T - mu * s_n) + c_n with s_n = sum_k g_k W_nk, c_n = sum_k b_k W_nk
//     (emulated on the checkpoint: tools/emulate_rawln.py, 1.5e-6 vs 1.1e-6 for normalise-then-split).  Row statistics
//     are accumulated by the epilogue that produced the rows (mean / M2 per 32-column block, combined exactly).  One A
//     operand therefore serves the self-attention Q/K/V projection AND the cross-attention K/V projection of a layer
//     (one op, N = 1280);
//   * the residual stream X is kept in fp32 next to its planes; the epilogue of proj / FFN2 adds it thread-per-row;
//   * the window-wide K / V of the pruned last layer leave as fp32 rows (F2_OUT_F32) for the newest-frame tail (k_tail);
//   * ops whose inputs were produced by the same CTA (attention after its projections) are separated by a CTA
//     barrier instead of a cluster barrier.
#pragma once

#include "common.cuh"
#include "gemm_tc.cuh"

namespace vapb {

enum F2Kind { F2_GEMM = 0, F2_ATTN = 1, F2_GATHER = 2 };
enum F2Out { F2_OUT_PLANES = 0, F2_OUT_X = 1, F2_OUT_F32 = 2 };
enum F2Side { F2_SIDE_NONE = 0, F2_SIDE_VAD = 1, F2_SIDE_GATHER_LAST = 2 };

struct F2Fields {                    // 128 bytes
    int kind;
    int K, N;                        // GEMM: K in {256, 768}, N total (multiple of 256)
    int out_mode;                    // F2Out
    int act;                         // 1 = exact-erf GELU
    int n_ln;                        // output columns [0, n_ln) get the LayerNorm correction
    int sibling;                     // ATTN: 1 = keys / values of the other channel
    int qcol, kcol, vcol;            // ATTN: first column of Q / K / V in their plane tensors (head h at + 64 h)
    int side;                        // F2Side, run by a spare warp during this op
    int cta_sync;                    // 1: the NEXT op only reads what this CTA wrote -> CTA barrier instead of cluster barrier
    int ld_out;                      // leading dimension (elements) of out_hi / out_lo
    int pad0;
    const float* ln_s;               // [n_ln]
    const float* ln_c;               // [n_ln]
    __nv_bfloat16* out_hi;           // planes out (tile-row layout); F2_OUT_X: the X planes
    __nv_bfloat16* out_lo;
    float* out_f;                    // F2_OUT_F32: columns [0, 512) -> out_f, [512, 1024) -> out_f2, rows (2b + c) * T + p, ld 512
    float* out_f2;
    const float* slopes;             // ATTN: ALiBi slopes [4]
    int pad1[4];
};
struct alignas(64) F2Op {
    // GEMM: m[0] / m[1] = A hi / lo planes (box 64 rows x 64: each CTA of a stream multicasts half a tile),
    //       m[2] / m[3] = W hi / lo (box 64 n x 64 k: each CTA of a column half multicasts half a tile)
    // ATTN: m[0] / m[1] = Q planes (box 128 rows x 64), m[2] / m[3] = K / V planes (box 64 rows x 64)
    // every op: m[4] / m[5] = output hi / lo planes, [sequence][position < T][cols], box 1 x 32 x 32 columns (SWIZZLE_64B half tiles);
    //           F2_OUT_X: the X planes, m[6] = the fp32 residual stream, same 3-D shape, box 1 x 32 x 32 floats;
    //           F2_OUT_F32: m[4] / m[5] = fp32 [sequence][position][512] tensors of out_f / out_f2 (box 32 x 32 x 1: positions >= T are clipped)
    CUtensorMap m[7];
    F2Fields f;
};
static_assert(sizeof(F2Fields) == 128 && sizeof(F2Op) == 1024, "F2Op layout");

struct Fused2Params {
    const F2Op* ops;
    int n_ops;
    int T;
    int B;                   // streams in this step (the launch pads an odd batch with a ghost stream on scratch rows)
    const float* ring;       // [max_streams][2][T][256]
    float* ring_w;           // same buffer (the downsample tail appends the newest frame)
    const float* ds_part;    // see fused_tf.cuh
    long long ds_stride;
    int ds_nsplit;
    const float* ds_lnw;
    const float* ds_lnb;
    float* e_out;
    const int* count;
    const int* ids;
    int* tvalid;
    float* Xf;               // [B * 128][256] fp32 residual stream, tile-row layout
    __nv_bfloat16* Xh;       // its planes
    __nv_bfloat16* Xl;
    float* stats;            // [B * 128][8][2]: mean, M2 of each 32-column block of X
    float* Xlast;            // [2B][256] newest frame per sequence (tail of the pruned layer)
    const float* va_w;
    const float* va_b;
    float* out;
    const IoPtrs* io;
    long long* dbg;          // optional clock64 stamps [n_ops + 1] of cluster 0 / CTA 0; fine stamps of op dbg_op at [40..52)
    int dbg_op;
};

size_t fused2_smem_bytes();
bool fused2_prepare(std::string& err);
cudaError_t launch_fused_tf2(const Fused2Params& p, int B, cudaStream_t st);

}  // namespace vapb
