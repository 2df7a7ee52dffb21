// Per-stream persistent transformer kernel ("stream kernel").
//
// One thread-block cluster of two CTAs owns ONE stereo stream for the whole transformer stack
// (ring gather -> ar_channel layer -> vad -> cross layers -> K/V of the pruned last layer): every
// dependency of the stack is local to a stream (rows of a sequence, plus the sibling channel for the
// cross attention, modules.py:276-283), so the ~30 ops run back to back inside one launch with a
// cluster barrier between them instead of ~35 dependent kernel launches.  The op list is a small
// "program" in global memory, interpreted by every cluster.
//
//   mode 0 (2T <= 128): the M tile is both channels of the stream (2T rows); CTA r computes N half r.
//   mode 1 (T <= 128):  the M tile is channel r of the stream (T rows); each CTA computes the full N.
#pragma once

#include "common.cuh"
#include "gemm_tc.cuh"

namespace vapb {

enum FOpKind { FOP_GEMM = 0, FOP_ATTN = 1, FOP_GATHER_RING = 2 };
enum FSide { FSIDE_NONE = 0, FSIDE_VAD = 1, FSIDE_GATHER_LAST = 2 };

struct FOpFields {                   // 128 bytes: copied to shared memory one op ahead
    int kind;
    // GEMM: C[rows, N] = act(LN?(A[rows, K]) W^T) + R      (R and C share ldc; R may alias C)
    int K, N, act, lda, ldc;
    // attention
    int sibling, ldq, ldk, ldv, ldo;
    int side;            // side task of the spare warp during this op: 0 none, 1 vad of the ar_channel output, 2 newest-frame gather
    const float* A;
    const float* ln_w;
    const float* ln_b;
    const float* R;
    float* C;
    const float* Q;
    const float* Kp;
    const float* V;
    float* O;
    const float* slopes;
};
struct alignas(64) FOp {
    CUtensorMap map_hi, map_lo;      // W planes, box {64 k, 64 n}, SWIZZLE_128B (GEMM only)
    FOpFields f;
};
static_assert(sizeof(FOpFields) == 128 && sizeof(FOp) == 384, "FOp layout");

struct FusedParams {
    const FOp* ops;
    int n_ops;
    int T;
    int mode;
    float* ring;             // [max_streams][2][T][256]
    // downsample tail folded into the ring gather (encoder_components.py:496-511): split-K partials of the downsample
    // GEMM [ds_nsplit][2B][256] -> sum -> LayerNorm -> GELU -> ring slot count % T (+ X row t-1, + e_out); null = the ring
    // already holds the newest frame
    const float* ds_part;
    long long ds_stride;
    int ds_nsplit;
    const float* ds_lnw;
    const float* ds_lnb;
    float* e_out;            // [2B][256] tap of the new embedding
    const int* count;
    const int* ids;
    int* tvalid;
    float* X;                // [2B*T][256]
    float* Xl;               // [2B][256]
    const float* va_w;
    const float* va_b;
    float* out;              // [B][6]
    const IoPtrs* io;        // != null: out = io->out (graph launches)
    long long* dbg;          // optional clock64 stamps: [n_ops + 1] of cluster 0 / CTA 0, fine stamps of op dbg_op at [40..52)
    int dbg_op;
};

size_t fused_smem_bytes();
bool fused_prepare(std::string& err);      // cudaFuncSetAttribute, once per device
// launches 2B CTAs (clusters of 2)
cudaError_t launch_fused_tf(const FusedParams& p, int B, cudaStream_t st);

}  // namespace vapb
