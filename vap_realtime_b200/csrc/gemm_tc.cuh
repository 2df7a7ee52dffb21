// tcgen05 (5th-gen tensor core) GEMM with bf16 hi/lo split operands ("bf16 x3").
//
//   C[M,N] = act(A[M,K] * W[N,K]^T + bias) + R        fp32 in, fp32 out
//
// fp32 parity with the reference needs more than one bf16/tf32 product
// (plain TF32 / BF16 miss the 1e-4 gate).  Every operand is split
// as x = hi + lo (both bf16) and three MMAs accumulate into ONE fp32 TMEM tile:
//   hi*hi + hi*lo + lo*hi       (the lo*lo term is below fp32 resolution)
//  * W planes are split once at vapb_create and fetched with TMA (128B swizzle).
//  * A stays fp32 in global memory; producer warps load it with coalesced 128-bit
//    loads, split it in registers and store the two planes straight into the
//    swizzled K-major smem layout the UMMA descriptors expect.  Because A rows are
//    addressed through a RowMap, the Conv1d layers run im2col-free on the
//    channels-last activations (overlapping rows, zero halo rows = padding).
#pragma once

#include "common.cuh"

#include <cuda.h>
#include <cuda_bf16.h>

#include <string>
#include <vector>

namespace vapb {

struct TcWeight {
    __nv_bfloat16* hi = nullptr;     // [N][K]
    __nv_bfloat16* lo = nullptr;     // [N][K]
    int N = 0, K = 0;
    bool has_tile[3] = {false, false, false};      // N tile widths 64 / 128 / 256
    CUtensorMap map_hi[3], map_lo[3];              // 2D {K, N}, box {64, tile}, SWIZZLE_128B
    CUtensorMap map_hi32, map_lo32;                // box {64, 32}: half tiles for the 2-CTA multicast
};

struct TcWorkspace {
    int force_bn = 0;                // 0 = pick the tile width from the grid size, else 64 / 128 / 256
    int use_k256 = 1;                // route K = 256 GEMMs to the resident-A kernel
    int cluster2 = 0;                // ... as 2-CTA clusters that multicast the W tiles (measured 2 % slower at B = 64: r01_l_option_ablation.log)
};

// Splits W (device fp32 [N][K]) into bf16 planes and encodes the TMA descriptors.
bool tc_prepare_weight(const float* W_dev, int N, int K, TcWeight& out, std::vector<void*>& allocs, std::string& err);
bool tc_prepare_workspace(TcWorkspace& ws, size_t max_a_elems, std::vector<void*>& allocs, std::string& err);

// 2D tensor map over a row-major bf16 tensor [rows][cols] (cols contiguous): box {64 cols, box_rows}, 128 B swizzle.
bool tc_encode_bf16_2d(CUtensorMap* map, void* ptr, size_t rows, size_t cols, int box_rows, std::string& err);
bool tc_encode_bf16_2d_half(CUtensorMap* map, void* ptr, size_t rows, size_t cols, int box_rows, std::string& err);

// fp32 maps for TMA STORES out of 128 B-swizzled staging tiles: box {32 floats, box_rows [, 1]}
bool tc_encode_f32_2d(CUtensorMap* map, void* ptr, size_t rows, size_t cols, int box_rows, std::string& err);
bool tc_encode_f32_3d(CUtensorMap* map, void* ptr, size_t d2, size_t d1, size_t cols, int box_d1, std::string& err, size_t d2_stride_rows = 0);
bool tc_encode_bf16_3d_half(CUtensorMap* map, void* ptr, size_t n_seq, size_t n_pos, size_t cols, size_t seq_stride_rows, std::string& err);

int tc_pick_ksplit(int M, int N, int K, int max_split);

// Enqueues the GEMM; returns the number of kernels launched.
int launch_gemm_tc(const GemmArgs& g, const TcWeight& w, TcWorkspace& ws, cudaStream_t st);

// Runs the tensor-core GEMM against the fp32 CUDA-core GEMM on random data.
int tc_selftest(int device, int variant, double* max_rel_err, std::string& report);

}  // namespace vapb
