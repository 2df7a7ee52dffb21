// CUDA-core (true fp32 FMA) kernels of the VAP streaming step.
//
// These are the precision-critical ops that must stay in fp32 (conv0, the LSTM
// recurrence) plus the row-wise / attention / head kernels, and an
// fp32 GEMM that serves as the exact-arithmetic mode of the library
// (option "gemm"=0) against which the tcgen05 bf16x3 path is validated.
//
// Reference semantics restated per kernel (file:line under the reference tree).
#include "common.cuh"

#include <cstdlib>

#include <math.h>

namespace vapb {

thread_local bool g_use_pdl = false;      // per thread: two handles may be driven from two threads
thread_local bool g_attn_rk = true;      // register-resident-K attention for T <= 64 (option "attn_rk")

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float gelu_erf(float x) {
    // nn.GELU() default = exact erf form (modules.py:9-21, encoder_components.py:506)
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// one 256-wide fp32 row spread over a warp: lane owns columns {4*lane..+3, 128+4*lane..+3}
__device__ __forceinline__ void load_row8(const float* p, int lane, float v[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p + lane * 4);
    const float4 b = *reinterpret_cast<const float4*>(p + 128 + lane * 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store_row8(float* p, int lane, const float v[8]) {
    *reinterpret_cast<float4*>(p + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 128 + lane * 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// -----------------------------------------------------------------------------------------
// conv0 (1 -> 256, k=10, s=5, p=3) + ChannelNorm + ReLU, channels-last output.
// encoder_components.py:83-84, 99 ; ChannelNorm :62-70 (unbiased variance, eps in rsqrt).
// One warp per output position; lane owns channels {4*lane..4*lane+3, 128+4*lane..+3}.
// -----------------------------------------------------------------------------------------
constexpr int kC0PosPerBlock = 56;

__global__ void __launch_bounds__(256) k_conv0_cn_relu(const float* __restrict__ audio, const IoPtrs* __restrict__ io, int S, int L0,
                                                       const float* __restrict__ w,    // [10][256] tap-major
                                                       const float* __restrict__ b,
                                                       const float* __restrict__ cnw,
                                                       const float* __restrict__ cnb,
                                                       float* __restrict__ out, RowMap omap) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_in[kC0PosPerBlock * 5 + 16];
    const int chunk = blockIdx.y;
    const int p0 = blockIdx.x * kC0PosPerBlock;
    const int npos = min(kC0PosPerBlock, L0 - p0);
    const int first = 5 * p0 - 3;
    const int need = 5 * (npos - 1) + 10;
    const float* a = (io ? io->audio : audio) + (size_t)chunk * S;
    for (int i = threadIdx.x; i < need; i += blockDim.x) {
        int s = first + i;
        s_in[i] = (s >= 0 && s < S) ? a[s] : 0.0f;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float wr[8][10], br[8], gw[8], gb[8];
#pragma unroll
    for (int k = 0; k < 10; ++k) {          // coalesced float4 loads of the tap-major weights
        float t[8];
        load_row8(w + k * 256, lane, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) wr[i][k] = t[i];
    }
    load_row8(b, lane, br);
    load_row8(cnw, lane, gw);
    load_row8(cnb, lane, gb);
    __syncthreads();
    for (int p = warp; p < npos; p += 8) {
        float x[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) x[k] = s_in[5 * p + k];
        float v[8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 10; ++k) acc = fmaf(wr[i][k], x[k], acc);
            v[i] = acc + br[i];
            s += v[i];
        }
        const float mean = warp_sum(s) * (1.0f / 256.0f);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float d = v[i] - mean;
            q = fmaf(d, d, q);
        }
        const float var = warp_sum(q) * (1.0f / 255.0f);
        const float rstd = 1.0f / sqrtf(var + kEps);
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaxf((v[i] - mean) * rstd * gw[i] + gb[i], 0.0f);
        float* dst = out + rowmap_off(omap, chunk * L0 + p0 + p);
        *reinterpret_cast<float4*>(dst + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(dst + 128 + lane * 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
}

void launch_conv0(const float* audio, const IoPtrs* io, int n_chunks, int S, int L0, const float* w, const float* b,
                  const float* cnw, const float* cnb, float* out, RowMap omap, cudaStream_t st) {
    dim3 grid((L0 + kC0PosPerBlock - 1) / kC0PosPerBlock, n_chunks);
    launch_k(k_conv0_cn_relu, grid, dim3(256), 0, st, audio, io, S, L0, w, b, cnw, cnb, out, omap);
}

// -----------------------------------------------------------------------------------------
// fp32 GEMM  C[M,N] = act(A[M,K] * W[N,K]^T + bias) + R     (both operands K-contiguous)
// 128x128x8 tiles, 256 threads, 8x8 register block, smem double buffering.
// N % 128 == 0, K % 8 == 0 (true for every shape on this path); M arbitrary.
// -----------------------------------------------------------------------------------------
constexpr int BM = 128, BN = 128, BK = 8, LDS = 132;

__global__ void __launch_bounds__(256) k_sgemm(GemmArgs g) {
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(16) float As[2][BK][LDS];
    __shared__ __align__(16) float Bs[2][BK][LDS];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int lr = tid >> 1, kq = (tid & 1) * 4;
    const int am = m0 + lr;
    const float* aptr = (am < g.M) ? (g.A + rowmap_off(g.amap, am) + kq) : nullptr;
    const float* wptr = g.W + (size_t)(n0 + lr) * g.K + kq;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 ra = aptr ? __ldg(reinterpret_cast<const float4*>(aptr)) : z4;
    float4 rw = __ldg(reinterpret_cast<const float4*>(wptr));
    As[0][kq + 0][lr] = ra.x; As[0][kq + 1][lr] = ra.y; As[0][kq + 2][lr] = ra.z; As[0][kq + 3][lr] = ra.w;
    Bs[0][kq + 0][lr] = rw.x; Bs[0][kq + 1][lr] = rw.y; Bs[0][kq + 2][lr] = rw.z; Bs[0][kq + 3][lr] = rw.w;
    __syncthreads();

    const int tx = tid & 15, ty = tid >> 4;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = g.K / BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            ra = aptr ? __ldg(reinterpret_cast<const float4*>(aptr + (size_t)(kt + 1) * BK)) : z4;
            rw = __ldg(reinterpret_cast<const float4*>(wptr + (size_t)(kt + 1) * BK));
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            const int nx = cur ^ 1;
            As[nx][kq + 0][lr] = ra.x; As[nx][kq + 1][lr] = ra.y; As[nx][kq + 2][lr] = ra.z; As[nx][kq + 3][lr] = ra.w;
            Bs[nx][kq + 0][lr] = rw.x; Bs[nx][kq + 1][lr] = rw.y; Bs[nx][kq + 2][lr] = rw.z; Bs[nx][kq + 3][lr] = rw.w;
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4)));
        if (m >= g.M) continue;
        float* crow = g.C + rowmap_off(g.cmap, m);
        const float* rrow = g.R ? (g.R + rowmap_off(g.rmap, m)) : nullptr;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const int n = n0 + jj * 64 + tx * 4;
            float v[4] = {acc[i][jj * 4 + 0], acc[i][jj * 4 + 1], acc[i][jj * 4 + 2], acc[i][jj * 4 + 3]};
            if (g.bias) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(g.bias + n));
                v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
            }
            if (g.act == 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = gelu_erf(v[q]);
            }
            if (rrow) {
                const float4 rr = *reinterpret_cast<const float4*>(rrow + n);
                v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
            }
            *reinterpret_cast<float4*>(crow + n) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// Small-problem variant: 64x64x32 tiles (4x more CTAs, 4x fewer, fatter k-iterations) for the
// latency-bound shapes (LSTM input projection M = 2B*5, anything with a handful of 128-row tiles).
constexpr int SM_ = 64, SN_ = 64, SK_ = 32, SLD = 68;

__global__ void __launch_bounds__(256) k_sgemm64(GemmArgs g) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // lets the LSTM kernel behind it start its weight fill early (no-op otherwise)
    pdl_wait();
    __shared__ __align__(16) float As[2][SK_][SLD];
    __shared__ __align__(16) float Bs[2][SK_][SLD];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * SM_, n0 = blockIdx.y * SN_;
    // loader: thread -> (row lr = tid / 4 (0..63), k quads kq = (tid % 4) * 4 and + 16)
    const int lr = tid >> 2, kq = (tid & 3) * 4;
    const int am = m0 + lr;
    const float* aptr = (am < g.M) ? (g.A + rowmap_off(g.amap, am) + kq) : nullptr;
    const float* wptr = g.W + (size_t)(n0 + lr) * g.K + kq;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // register prefetch two k-tiles ahead (an L2 / HBM round trip is longer than one tile of FMAs), smem double buffer
    struct Regs { float4 a0, a1, w0, w1; };
    auto gload = [&](int kt, Regs& r) {
        const size_t o = (size_t)kt * SK_;
        r.a0 = aptr ? __ldg(reinterpret_cast<const float4*>(aptr + o)) : z4;
        r.a1 = aptr ? __ldg(reinterpret_cast<const float4*>(aptr + o + 16)) : z4;
        r.w0 = __ldg(reinterpret_cast<const float4*>(wptr + o));
        r.w1 = __ldg(reinterpret_cast<const float4*>(wptr + o + 16));
    };
    auto sstore = [&](int buf, const Regs& r) {
        As[buf][kq + 0][lr] = r.a0.x; As[buf][kq + 1][lr] = r.a0.y; As[buf][kq + 2][lr] = r.a0.z; As[buf][kq + 3][lr] = r.a0.w;
        As[buf][kq + 16][lr] = r.a1.x; As[buf][kq + 17][lr] = r.a1.y; As[buf][kq + 18][lr] = r.a1.z; As[buf][kq + 19][lr] = r.a1.w;
        Bs[buf][kq + 0][lr] = r.w0.x; Bs[buf][kq + 1][lr] = r.w0.y; Bs[buf][kq + 2][lr] = r.w0.z; Bs[buf][kq + 3][lr] = r.w0.w;
        Bs[buf][kq + 16][lr] = r.w1.x; Bs[buf][kq + 17][lr] = r.w1.y; Bs[buf][kq + 18][lr] = r.w1.z; Bs[buf][kq + 19][lr] = r.w1.w;
    };
    const int nk = g.K / SK_;
    Regs r0, r1;
    gload(0, r0);
    if (nk > 1) gload(1, r1);
    sstore(0, r0);
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads, 4 x 4 outputs each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    auto compute = [&](int cur) {
#pragma unroll
        for (int k = 0; k < SK_; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    };
    // two k-tiles per trip so that the register sets keep their names (r1 holds tile kt + 1, r0 gets tile kt + 2, ...)
    for (int kt = 0; kt < nk; kt += 2) {
        if (kt + 2 < nk) gload(kt + 2, r0);
        compute(0);
        if (kt + 1 < nk) sstore(1, r1);
        __syncthreads();
        if (kt + 1 < nk) {
            if (kt + 3 < nk) gload(kt + 3, r1);
            compute(1);
            if (kt + 2 < nk) sstore(0, r0);
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
        const int n = n0 + tx * 4;
        float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
        if (g.bias) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(g.bias + n));
            v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
        }
        if (g.act == 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = gelu_erf(v[q]);
        }
        if (g.R) {
            const float4 rr = *reinterpret_cast<const float4*>(g.R + rowmap_off(g.rmap, m) + n);
            v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
        }
        *reinterpret_cast<float4*>(g.C + rowmap_off(g.cmap, m) + n) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

void launch_sgemm(const GemmArgs& g, cudaStream_t st) {
    const int big_ctas = ((g.M + BM - 1) / BM) * (g.N / BN);
    if (big_ctas < 100 && g.K % SK_ == 0) {
        dim3 grid((g.M + SM_ - 1) / SM_, g.N / SN_);
        launch_k(k_sgemm64, grid, dim3(256), 0, st, g);
    } else {
        dim3 grid((g.M + BM - 1) / BM, g.N / BN);
        launch_k(k_sgemm, grid, dim3(256), 0, st, g);
    }
}

// -----------------------------------------------------------------------------------------
// ChannelNorm + ReLU in place on rows of 256 (encoder_components.py:62-70, 99-103).
// LayerNorm(256) (+ optional GELU) (modules.py:242-243, 268; encoder_components.py:408-428).
// One warp per row; lane owns columns {4*lane.., 128+4*lane..}.
// -----------------------------------------------------------------------------------------
// normalise v[8] (one 256-wide row spread over a warp); denom = 255 (unbiased) or 256 (biased)
__device__ __forceinline__ void warp_norm8(float v[8], float denom_inv, const float* w, const float* b, int lane) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    const float mean = warp_sum(s) * (1.0f / 256.0f);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float d = v[i] - mean;
        q = fmaf(d, d, q);
    }
    const float var = warp_sum(q) * denom_inv;
    const float rstd = 1.0f / sqrtf(var + kEps);
    float ww[8], bb[8];
    load_row8(w, lane, ww);
    load_row8(b, lane, bb);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * ww[i] + bb[i];
}

// With split-K partials ([nsplit][M][256], plain layout) the rows are summed here first.
__global__ void __launch_bounds__(256) k_cn_relu(float* X, RowMap map, int M, const float* __restrict__ w,
                                                 const float* __restrict__ b, const float* __restrict__ partials, int nsplit,
                                                 long long split_stride) {
    pdl_trigger();
    pdl_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    float* p = X + rowmap_off(map, row);
    float v[8];
    if (nsplit > 0) {
        load_row8(partials + (size_t)row * kD, lane, v);
        for (int z = 1; z < nsplit; ++z) {
            float u[8];
            load_row8(partials + (size_t)z * split_stride + (size_t)row * kD, lane, u);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += u[i];
        }
    } else {
        load_row8(p, lane, v);
    }
    warp_norm8(v, 1.0f / 255.0f, w, b, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
    store_row8(p, lane, v);
}
void launch_cn_relu(float* X, RowMap map, int M, const float* w, const float* b, cudaStream_t st, const float* partials,
                    int nsplit, long long split_stride) {
    launch_k(k_cn_relu, dim3((M + 7) / 8), dim3(256), 0, st, X, map, M, w, b, partials, nsplit, split_stride);
}

__global__ void __launch_bounds__(256) k_layernorm(const float* __restrict__ X, RowMap xmap, float* __restrict__ Y,
                                                   RowMap ymap, int M, const float* __restrict__ w,
                                                   const float* __restrict__ b, int gelu) {
    pdl_trigger();
    pdl_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    float v[8];
    load_row8(X + rowmap_off(xmap, row), lane, v);
    warp_norm8(v, 1.0f / 256.0f, w, b, lane);
    if (gelu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
    }
    store_row8(Y + rowmap_off(ymap, row), lane, v);
}
void launch_layernorm(const float* X, RowMap xmap, float* Y, RowMap ymap, int M, const float* w,
                      const float* b, int gelu, cudaStream_t st) {
    launch_k(k_layernorm, dim3((M + 7) / 8), dim3(256), 0, st, X, xmap, Y, ymap, M, w, b, gelu);
}

// -----------------------------------------------------------------------------------------
// LSTM state staging and cell (encoder_components.py:120-123, 140-153; gate order i,f,g,o).
// State arrays are [max_streams][2][256]; work arrays are [2B][256] in batch order.
// -----------------------------------------------------------------------------------------
__global__ void k_gather_state(const float* __restrict__ hS, const float* __restrict__ cS,
                               const int* __restrict__ ids, float* hW, float* cW, int B) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // float4 index over [2B][64]
    if (i >= B * 2 * 64) return;
    const int n = i / 64, q = i % 64;
    const int b = n >> 1, ch = n & 1;
    const size_t src = ((size_t)ids[b] * 2 + ch) * 64 + q;
    reinterpret_cast<float4*>(hW)[i] = reinterpret_cast<const float4*>(hS)[src];
    reinterpret_cast<float4*>(cW)[i] = reinterpret_cast<const float4*>(cS)[src];
}
void launch_gather_state(const float* hS, const float* cS, const int* ids, float* hW, float* cW, int B,
                         cudaStream_t st) {
    const int n = B * 2 * 64;
    launch_k(k_gather_state, dim3((n + 255) / 256), dim3(256), 0, st, hS, cS, ids, hW, cW, B);
}
__global__ void k_scatter_state(float* hS, float* cS, const int* __restrict__ ids, const float* __restrict__ hW,
                                const float* __restrict__ cW, int B) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 2 * 64) return;
    const int n = i / 64, q = i % 64;
    const int b = n >> 1, ch = n & 1;
    const size_t dst = ((size_t)ids[b] * 2 + ch) * 64 + q;
    reinterpret_cast<float4*>(hS)[dst] = reinterpret_cast<const float4*>(hW)[i];
    reinterpret_cast<float4*>(cS)[dst] = reinterpret_cast<const float4*>(cW)[i];
}
void launch_scatter_state(float* hS, float* cS, const int* ids, const float* hW, const float* cW, int B,
                          cudaStream_t st) {
    const int n = B * 2 * 64;
    launch_k(k_scatter_state, dim3((n + 255) / 256), dim3(256), 0, st, hS, cS, ids, hW, cW, B);
}

// G [n_rows][1024] = W_ih x_t + b_ih + W_hh h + b_hh (already summed by the GEMMs)
__global__ void k_lstm_cell(const float* __restrict__ G, float* hW, float* cW, float* __restrict__ Y, int n_rows,
                            int n_steps, int step) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * kD) return;
    const int n = i / kD, d = i % kD;
    const float* g = G + (size_t)n * 4 * kD;
    const float gi = g[d], gf = g[kD + d], gg = g[2 * kD + d], go = g[3 * kD + d];
    const float c = sigmoidf_(gf) * cW[i] + sigmoidf_(gi) * tanhf(gg);
    const float h = sigmoidf_(go) * tanhf(c);
    cW[i] = c;
    hW[i] = h;
    Y[((size_t)n * n_steps + step) * kD + d] = h;
}
void launch_lstm_cell(const float* G, float* hW, float* cW, float* Y, int n_rows, int n_steps, int step,
                      cudaStream_t st) {
    const int n = n_rows * kD;
    launch_k(k_lstm_cell, dim3((n + 255) / 256), dim3(256), 0, st, G, hW, cW, Y, n_rows, n_steps, step);
}

// -----------------------------------------------------------------------------------------
// Fused LSTM recurrence (encoder_components.py:120-123, 140-153): all n_steps of
//   g = Gx[:, t] + W_hh h ;  c = sig(f) c + sig(i) tanh(g~) ;  h = sig(o) tanh(c)
// in ONE kernel, true fp32.  A thread-block cluster of 8 CTAs owns a tile of kLstmRT rows
// (channel-chunks); CTA r owns hidden units [32r, 32r+32) = 128 gate columns whose W_hh slice
// (128 KB) stays resident in shared memory for every step.  After each step the new h slices
// are exchanged through distributed shared memory (st to the 8 peers) + one cluster barrier.
// Also replaces the gather/scatter of the per-stream (h, c) state.
// -----------------------------------------------------------------------------------------
constexpr int kLstmWPitch = 128;      // W_hh slice as float4 [k / 4][128 gate columns]: one 128-bit conflict-free read per 4 k
constexpr int kLstmGPitch = 132;
template <int RT>
constexpr size_t lstm_smem() { return (size_t)(256 * kLstmWPitch + 2 * RT * kD + RT * kLstmGPitch + RT * 32) * sizeof(float); }

template <int kLstmRT>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(256)
k_lstm_recurrent(const float* __restrict__ Gx,      // [NC][n_steps][1024]  W_ih z + b_ih + b_hh
                 const float* __restrict__ Whh,     // [1024][256]
                 float* hS, float* cS, const int* __restrict__ ids,
                 float* __restrict__ Y,             // [NC][n_steps][256]
                 int NC, int n_steps) {
    pdl_trigger();
    extern __shared__ float lsm[];
    float* sW = lsm;                                   // float4 [64 k-quads][128]  (col = gate*32 + unit)
    float* sH = sW + 256 * kLstmWPitch;                // [2][RT][256]
    float* sG = sH + 2 * kLstmRT * kD;                 // [RT][132]
    float* sC = sG + kLstmRT * kLstmGPitch;            // [RT][32]
    unsigned crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int tid = threadIdx.x;
    const int row0 = (blockIdx.x >> 3) * kLstmRT;

    // W_hh slice -> smem, transposed to [k][col]: 128 columns x 64 float4 along k, 8 loads in flight per thread
#pragma unroll 1
    for (int base = 0; base < 128 * 64; base += 256 * 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = base + u * 256 + tid;              // (k-quad, col); consecutive lanes -> consecutive columns
            const int col = idx & 127, kq = idx >> 7;
            const int grow = (col >> 5) * kD + 32 * (int)crank + (col & 31);
            v[u] = __ldg(reinterpret_cast<const float4*>(Whh + (size_t)grow * kD) + kq);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) reinterpret_cast<float4*>(sW)[base + u * 256 + tid] = v[u];
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");     // programmatic dependent launch: the weight fill above overlaps the
                                                            // input-projection GEMM in front of this kernel; everything below depends on it
    // initial h (all 256 units of the tile's rows) and c (own 32 units)
    for (int i = tid; i < kLstmRT * kD; i += 256) {
        const int r = i >> 8, k = i & 255, n = row0 + r;
        float v = 0.f;
        if (n < NC) v = hS[((size_t)ids[n >> 1] * 2 + (n & 1)) * kD + k];
        sH[i] = v;
    }
    for (int i = tid; i < kLstmRT * 32; i += 256) {
        const int r = i >> 5, u = i & 31, n = row0 + r;
        float v = 0.f;
        if (n < NC) v = cS[((size_t)ids[n >> 1] * 2 + (n & 1)) * kD + 32 * crank + u];
        sC[i] = v;
    }
    // every CTA of the cluster is running and initialised before the first DSMEM store
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");

    constexpr int RH = kLstmRT / 2;                     // rows per thread in the GEMV phase
    constexpr int RC = (kLstmRT + 7) / 8;               // rows per thread in the cell phase (rows cr + 8 * rr < RT)
    const int col = tid & 127, rhalf = tid >> 7;       // thread -> one gate column, half of the rows
    const int gcol = (col >> 5) * kD + 32 * (int)crank + (col & 31);
    const int cu = tid & 31, cr = tid >> 5;            // cell phase: unit cu, rows cr + 8 * rr
    uint32_t sH_u32 = (uint32_t)__cvta_generic_to_shared(sH);
    float h_last[RC];
#pragma unroll
    for (int rr = 0; rr < RC; ++rr) h_last[rr] = 0.f;

    for (int t = 0; t < n_steps; ++t) {
        const float* hcur = sH + (t & 1) * kLstmRT * kD;
        float gx[RH];
#pragma unroll
        for (int r = 0; r < RH; ++r) {
            const int n = row0 + rhalf * RH + r;
            gx[r] = (n < NC) ? __ldg(Gx + ((size_t)n * n_steps + t) * 4 * kD + gcol) : 0.f;
        }
        float acc[RH];
#pragma unroll
        for (int r = 0; r < RH; ++r) acc[r] = 0.f;
#pragma unroll 4
        for (int k = 0; k < kD; k += 4) {
            const float4 w4 = reinterpret_cast<const float4*>(sW)[(k >> 2) * 128 + col];
            const float w0 = w4.x, w1 = w4.y, w2 = w4.z, w3 = w4.w;
#pragma unroll
            for (int r = 0; r < RH; ++r) {
                const float4 hv = *reinterpret_cast<const float4*>(hcur + (rhalf * RH + r) * kD + k);
                acc[r] = fmaf(w0, hv.x, acc[r]);
                acc[r] = fmaf(w1, hv.y, acc[r]);
                acc[r] = fmaf(w2, hv.z, acc[r]);
                acc[r] = fmaf(w3, hv.w, acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < RH; ++r) sG[(rhalf * RH + r) * kLstmGPitch + col] = acc[r] + gx[r];
        __syncthreads();
        // cell update for (unit cu, rows cr / cr+8); broadcast the new h to all 8 CTAs of the cluster
        const uint32_t nxt_off = (uint32_t)(((t + 1) & 1) * kLstmRT * kD) * 4u;
#pragma unroll
        for (int rr = 0; rr < RC; ++rr) {
            const int r = cr + 8 * rr;
            if (r >= kLstmRT) continue;
            const float* gr = sG + r * kLstmGPitch;
            const float gi = gr[cu], gf = gr[32 + cu], gg = gr[64 + cu], go = gr[96 + cu];
            const float c = sigmoidf_(gf) * sC[r * 32 + cu] + sigmoidf_(gi) * tanhf(gg);
            const float h = sigmoidf_(go) * tanhf(c);
            sC[r * 32 + cu] = c;
            h_last[rr] = h;
            const int n = row0 + r;
            if (n < NC) Y[((size_t)n * n_steps + t) * kD + 32 * crank + cu] = h;
            const uint32_t local = sH_u32 + nxt_off + (uint32_t)(r * kD + 32 * (int)crank + cu) * 4u;
#pragma unroll
            for (unsigned peer = 0; peer < 8; ++peer) {
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(peer));
                asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(h) : "memory");
            }
        }
        // release our DSMEM stores / acquire the peers' before anyone reads the next h buffer
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    // persist (h, c) of our 32 units
#pragma unroll
    for (int rr = 0; rr < RC; ++rr) {
        const int r = cr + 8 * rr, n = row0 + r;
        if (r < kLstmRT && n < NC) {
            const size_t o = ((size_t)ids[n >> 1] * 2 + (n & 1)) * kD + 32 * crank + cu;
            hS[o] = h_last[rr];
            cS[o] = sC[r * 32 + cu];
        }
    }
}

template <int RT>
static int lstm_max_clusters() {          // co-resident clusters of 8 CTAs on this device (GPCs with < 16 free SMs hold one, not two)
    cudaFuncSetAttribute(k_lstm_recurrent<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lstm_smem<RT>());
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(8 * 32);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = lstm_smem<RT>();
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k_lstm_recurrent<RT>, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = 8;
    }
    return n;
}

void launch_lstm_recurrent(const float* Gx, const float* Whh, float* hS, float* cS, const int* ids, float* Y, int NC,
                           int n_steps, cudaStream_t st) {
    // Row tile = the smallest of {8, 10, 12, 16} rows whose clusters all fit ONE wave: a second wave doubles the
    // kernel, and 16 clusters of 8 CTAs do not fit every B200 (15 on the dies measured: 128 chunks -> 10-row tiles,
    // 13 clusters on 104 SMs instead of 8 clusters on 64).  Above 16 rows per resident cluster: 16-row tiles, several waves.
    static OncePerDevice once;
    static int max_cl[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (once.first()) {
        const int a = lstm_max_clusters<8>(), b = lstm_max_clusters<10>(), c = lstm_max_clusters<12>(), d = lstm_max_clusters<16>();
        max_cl[dev] = a < b ? a : b;
        max_cl[dev] = max_cl[dev] < c ? max_cl[dev] : c;
        max_cl[dev] = max_cl[dev] < d ? max_cl[dev] : d;
    }
    static const int force_rt = getenv("VAPB_LSTM_RT") ? atoi(getenv("VAPB_LSTM_RT")) : 0;      // experiment knob
    int rt = 16;
    const int cands[4] = {8, 10, 12, 16};
    for (int i = 3; i >= 0; --i)
        if ((NC + cands[i] - 1) / cands[i] <= max_cl[dev]) rt = cands[i];
    if (force_rt) rt = force_rt;
    static const int use_pdl = getenv("VAPB_LSTM_PDL") ? atoi(getenv("VAPB_LSTM_PDL")) : 1;     // experiment knob
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cfg.attrs = at;
    cfg.numAttrs = use_pdl ? 1 : 0;
#define VAPB_LSTM_LAUNCH(RT_)                                                                        \
    do {                                                                                             \
        cfg.gridDim = dim3(((NC + RT_ - 1) / RT_) * 8);                                              \
        cfg.dynamicSmemBytes = lstm_smem<RT_>();                                                     \
        cudaLaunchKernelEx(&cfg, k_lstm_recurrent<RT_>, Gx, Whh, hS, cS, ids, Y, NC, n_steps);        \
    } while (0)
    if (rt == 8) VAPB_LSTM_LAUNCH(8);
    else if (rt == 10) VAPB_LSTM_LAUNCH(10);
    else if (rt == 12) VAPB_LSTM_LAUNCH(12);
    else VAPB_LSTM_LAUNCH(16);
#undef VAPB_LSTM_LAUNCH
}

// -----------------------------------------------------------------------------------------
// Downsample tail: LayerNorm + GELU, then append to the per-stream ring
// (encoder_components.py:496-511; vap_main.py:274-280).  The frame counter is
// advanced by the head kernel at the end of the step, so every kernel of a step
// sees count = frames BEFORE this step.
// ring layout: [max_streams][2][T][256]; slot of the new frame = count % T.
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ln_gelu_ring(const float* __restrict__ X, int B, const float* __restrict__ w,
                                                      const float* __restrict__ b, float* ring,
                                                      const int* __restrict__ count, const int* __restrict__ ids,
                                                      int T, float* e_out, int nsplit, long long split_stride) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // the stream kernel behind it sets up (TMEM, barriers, first W tiles) early
    pdl_wait();
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= 2 * B) return;
    const int lane = threadIdx.x & 31;
    float v[8];
    load_row8(X + (size_t)n * kD, lane, v);
    for (int z = 1; z < nsplit; ++z) {          // split-K partials of the downsample GEMM
        float u[8];
        load_row8(X + (size_t)z * split_stride + (size_t)n * kD, lane, u);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += u[i];
    }
    warp_norm8(v, 1.0f / 256.0f, w, b, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
    if (ring) {
        const int id = ids[n >> 1], ch = n & 1;
        const int slot = count[id] % T;
        store_row8(ring + (((size_t)id * 2 + ch) * T + slot) * kD, lane, v);
    }
    if (e_out) store_row8(e_out + (size_t)n * kD, lane, v);
}
void launch_ln_gelu_ring(const float* X, int B, const float* w, const float* b, float* ring, const int* count,
                         const int* ids, int T, float* e_out, cudaStream_t st, int nsplit, long long split_stride) {
    launch_k(k_ln_gelu_ring, dim3((2 * B + 7) / 8), dim3(256), 0, st, X, B, w, b, ring, count, ids, T, e_out, nsplit,
             split_stride);
}

// X[(n*T + j)] = ring row of logical position j (oldest first) for j < t, zero rows above.
// (torch.cat of the context list, vap_main.py:282-283.)
__global__ void __launch_bounds__(64) k_gather_ring(const float* __restrict__ ring, const int* __restrict__ count,
                                                    const int* __restrict__ ids, float* __restrict__ X,
                                                    int* __restrict__ tvalid, int B, int T) {
    pdl_trigger();
    pdl_wait();
    const int j = blockIdx.x, n = blockIdx.y;
    const int b = n >> 1, ch = n & 1;
    const int id = ids[b];
    const int cnt = count[id] + 1;                 // frames including the one appended this step
    const int t = cnt < T ? cnt : T;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < t) {
        const int slot = (cnt - t + j) % T;
        v = reinterpret_cast<const float4*>(ring + (((size_t)id * 2 + ch) * T + slot) * kD)[threadIdx.x];
    }
    reinterpret_cast<float4*>(X + ((size_t)n * T + j) * kD)[threadIdx.x] = v;
    if (j == 0 && ch == 0 && threadIdx.x == 0) tvalid[b] = t;
}
void launch_gather_ring(const float* ring, const int* count, const int* ids, float* X, int* tvalid, int B, int T,
                        cudaStream_t st) {
    dim3 grid(T, 2 * B);
    launch_k(k_gather_ring, grid, dim3(64), 0, st, ring, count, ids, X, tvalid, B, T);
}

// -----------------------------------------------------------------------------------------
// Causal ALiBi attention for one (sequence, head) per CTA (modules.py:82-110, 170-212):
//   S[i][j] = (q_i . k_j) / 16 + m_h * j   (j <= i),  P = softmax_j S,  O = P V
// scale = 1/sqrt(dim) = 1/16 (modules.py:52).  T <= 128.
// -----------------------------------------------------------------------------------------
constexpr int kAttnWarps = 8;

__global__ void __launch_bounds__(kAttnWarps * 32) k_attention(AttnArgs a) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float smem[];
    const int T = a.T;
    float* sK = smem;                       // [T][65] (row pitch 65 floats: conflict-free q.k dots)
    float* sV = sK + ((T * 65 + 3) & ~3);   // [T][64], 16-byte aligned for float4 stores
    float* sQ = sV + T * 64;                // [warps][64]
    float* sP = sQ + kAttnWarps * 64;       // [warps][128]
    const int n = blockIdx.x, h = blockIdx.y;
    const int t = a.tvalid[n >> 1];
    const int kvn = a.sibling ? (n ^ 1) : n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // layer-0 Q/K/V cache: rows live in the stream's ring, oldest frame at slot (frames - t) % T
    long long ring_base = -1;
    int ring_first = 0;
    if (a.ring_ids) {
        const int id = a.ring_ids[n >> 1];
        ring_base = ((long long)id * 2 + (n & 1)) * T;
        ring_first = (a.ring_count[id] + 1 - t) % T;
    }
    auto in_row = [&](int seq, int pos) -> size_t {
        return ring_base >= 0 ? (size_t)(ring_base + (ring_first + pos) % T) : (size_t)seq * T + pos;
    };
    for (int i = threadIdx.x; i < t * 16; i += blockDim.x) {
        const int j = i >> 4, q = (i & 15) * 4;
        const size_t r = in_row(kvn, j);
        const float4 kk = *reinterpret_cast<const float4*>(a.K + r * a.ldk + h * 64 + q);
        const float4 vv = *reinterpret_cast<const float4*>(a.V + r * a.ldv + h * 64 + q);
        float* dk = sK + j * 65 + q;
        dk[0] = kk.x; dk[1] = kk.y; dk[2] = kk.z; dk[3] = kk.w;
        *reinterpret_cast<float4*>(sV + j * 64 + q) = vv;
    }
    __syncthreads();
    const float slope = a.slopes[h];
    float* q = sQ + warp * 64;
    float* p = sP + warp * 128;
    for (int i = warp; i < T; i += kAttnWarps) {
        float* orow = a.O + ((size_t)n * T + i) * a.ldo + h * 64;
        if (i >= t) {                       // rows beyond the valid window: defined zeros
            orow[lane] = 0.f;
            orow[lane + 32] = 0.f;
            continue;
        }
        const float* qrow = a.Q + in_row(n, i) * a.ldq + h * 64;
        q[lane] = qrow[lane];
        q[lane + 32] = qrow[lane + 32];
        __syncwarp();
        float s[4];
        float mx = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = lane + 32 * jj;
            s[jj] = -INFINITY;
            if (j <= i) {
                const float* kr = sK + j * 65;
                float acc = 0.f;
#pragma unroll 16
                for (int d = 0; d < 64; ++d) acc = fmaf(q[d], kr[d], acc);
                s[jj] = acc * 0.0625f + slope * (float)j;
            }
            mx = fmaxf(mx, s[jj]);
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = lane + 32 * jj;
            const float e = (j <= i) ? expf(s[jj] - mx) : 0.f;
            s[jj] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = lane + 32 * jj;
            if (j < 128) p[j] = s[jj] * inv;
        }
        __syncwarp();
        float o0 = 0.f, o1 = 0.f;
        for (int j = 0; j <= i; ++j) {
            const float pj = p[j];
            o0 = fmaf(pj, sV[j * 64 + lane], o0);
            o1 = fmaf(pj, sV[j * 64 + lane + 32], o1);
        }
        orow[lane] = o0;
        orow[lane + 32] = o1;
        __syncwarp();
    }
}
// -----------------------------------------------------------------------------------------
// Attention for windows of at most 64 frames (the 2.5 s / 3 s models): the K rows of the
// (sequence, head) live in REGISTERS (lane j holds keys j and j+32), four query rows are processed
// per pass so that each broadcast q / p / V read from shared memory feeds 8..32 FMAs.
// Same arithmetic as k_attention (fp32 FMA, exact softmax).
// -----------------------------------------------------------------------------------------
template <int KC>
__global__ void __launch_bounds__(128) k_attention_rk(AttnArgs a) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float smem[];
    const int T = a.T;
    float* sV = smem;                          // [T][64]
    float* sQ = sV + T * 64;                   // [4 warps][4 rows][64]
    float* sP = sQ + 4 * 4 * 64;               // [4 warps][64 keys][4 rows]
    const int n = blockIdx.x, h = blockIdx.y;
    const int t = a.tvalid[n >> 1];
    const int kvn = a.sibling ? (n ^ 1) : n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < t * 16; i += blockDim.x) {
        const int j = i >> 4, q = (i & 15) * 4;
        *reinterpret_cast<float4*>(sV + j * 64 + q) =
            *reinterpret_cast<const float4*>(a.V + ((size_t)kvn * T + j) * a.ldv + h * 64 + q);
    }
    float kreg[KC][64];
#pragma unroll
    for (int c = 0; c < KC; ++c) {
        const int j = lane + 32 * c;
        if (j < t) {
            const float4* kp = reinterpret_cast<const float4*>(a.K + ((size_t)kvn * T + j) * a.ldk + h * 64);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float4 v = kp[i];
                kreg[c][4 * i] = v.x; kreg[c][4 * i + 1] = v.y; kreg[c][4 * i + 2] = v.z; kreg[c][4 * i + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 64; ++i) kreg[c][i] = 0.f;
        }
    }
    __syncthreads();
    const float slope = a.slopes[h];
    float* q = sQ + warp * 256;
    float* p = sP + warp * 256;
    for (int i0 = warp * 4; i0 < T; i0 += 16) {
        {   // four query rows -> smem (256 B per row, coalesced)
            const int r = lane >> 3, c8 = (lane & 7) * 8;
            const int i = min(i0 + r, T - 1);
            const float4* qp = reinterpret_cast<const float4*>(a.Q + ((size_t)n * T + i) * a.ldq + h * 64 + c8);
            *reinterpret_cast<float4*>(q + r * 64 + c8) = qp[0];
            *reinterpret_cast<float4*>(q + r * 64 + c8 + 4) = qp[1];
        }
        __syncwarp();
        float acc[KC][4];
#pragma unroll
        for (int c = 0; c < KC; ++c)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[c][r] = 0.f;
#pragma unroll
        for (int d = 0; d < 64; d += 4) {
            float4 q4[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) q4[r] = *reinterpret_cast<const float4*>(q + r * 64 + d);
#pragma unroll
            for (int c = 0; c < KC; ++c)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    acc[c][r] = fmaf(q4[r].x, kreg[c][d], acc[c][r]);
                    acc[c][r] = fmaf(q4[r].y, kreg[c][d + 1], acc[c][r]);
                    acc[c][r] = fmaf(q4[r].z, kreg[c][d + 2], acc[c][r]);
                    acc[c][r] = fmaf(q4[r].w, kreg[c][d + 3], acc[c][r]);
                }
        }
        float pr[KC][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = i0 + r;
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < KC; ++c) {
                const int j = lane + 32 * c;
                const float sc = (j <= i && j < t) ? (acc[c][r] * 0.0625f + slope * (float)j) : -INFINITY;
                pr[c][r] = sc;
                mx = fmaxf(mx, sc);
            }
            mx = warp_max(mx);
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < KC; ++c) {
                const float e = (pr[c][r] == -INFINITY) ? 0.f : expf(pr[c][r] - mx);
                pr[c][r] = e;
                sum += e;
            }
            sum = warp_sum(sum);
            const float inv = (i < t) ? (1.0f / sum) : 0.f;      // rows beyond the window produce zeros
#pragma unroll
            for (int c = 0; c < KC; ++c) pr[c][r] = (i < t) ? pr[c][r] * inv : 0.f;
        }
#pragma unroll
        for (int c = 0; c < KC; ++c)
            *reinterpret_cast<float4*>(p + (lane + 32 * c) * 4) = make_float4(pr[c][0], pr[c][1], pr[c][2], pr[c][3]);
        __syncwarp();
        float o[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r) o[r][0] = o[r][1] = 0.f;
        const int jmax = min(min(i0 + 3, T - 1), t - 1);
        for (int j = 0; j <= jmax; ++j) {
            const float4 p4 = *reinterpret_cast<const float4*>(p + j * 4);
            const float v0 = sV[j * 64 + lane], v1 = sV[j * 64 + lane + 32];
            o[0][0] = fmaf(p4.x, v0, o[0][0]); o[0][1] = fmaf(p4.x, v1, o[0][1]);
            o[1][0] = fmaf(p4.y, v0, o[1][0]); o[1][1] = fmaf(p4.y, v1, o[1][1]);
            o[2][0] = fmaf(p4.z, v0, o[2][0]); o[2][1] = fmaf(p4.z, v1, o[2][1]);
            o[3][0] = fmaf(p4.w, v0, o[3][0]); o[3][1] = fmaf(p4.w, v1, o[3][1]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = i0 + r;
            if (i < T) {
                float* orow = a.O + ((size_t)n * T + i) * a.ldo + h * 64;
                orow[lane] = o[r][0];
                orow[lane + 32] = o[r][1];
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) k_qkv_append(const float* __restrict__ qkv_new, float* __restrict__ qkv_ring, const int* __restrict__ count,
                                                    const int* __restrict__ ids, int B, int T) {
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.x;                  // sequence 2b + ch
    const int id = ids[n >> 1];
    const int slot = count[id] % T;            // the counter is advanced at the end of the step
    const float4* src = reinterpret_cast<const float4*>(qkv_new + (size_t)n * 3 * kD);
    float4* dst = reinterpret_cast<float4*>(qkv_ring + (((size_t)id * 2 + (n & 1)) * T + slot) * 3 * kD);
    if (threadIdx.x < 3 * kD / 4) dst[threadIdx.x] = src[threadIdx.x];
}
void launch_qkv_append(const float* qkv_new, float* qkv_ring, const int* count, const int* ids, int B, int T, cudaStream_t st) {
    launch_k(k_qkv_append, dim3(2 * B), dim3(256), 0, st, qkv_new, qkv_ring, count, ids, B, T);
}

void launch_attention(const AttnArgs& a, cudaStream_t st) {
    if (a.T <= 64 && g_attn_rk && !a.ring_ids) {
        const size_t sm = (size_t)(a.T * 64 + 4 * 256 + 4 * 256) * sizeof(float);
        dim3 grid(a.n_seq, kHeads);
        if (a.T <= 32) launch_k(k_attention_rk<1>, grid, dim3(128), sm, st, a);
        else launch_k(k_attention_rk<2>, grid, dim3(128), sm, st, a);
        return;
    }
    const size_t smem = (size_t)(((a.T * 65 + 3) & ~3) + a.T * 64 + kAttnWarps * 64 + kAttnWarps * 128) * sizeof(float);
    static OncePerDevice once;
    if (once.first()) cudaFuncSetAttribute(k_attention, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    dim3 grid(a.n_seq, kHeads);
    launch_k(k_attention, grid, dim3(kAttnWarps * 32), smem, st, a);
}

// -----------------------------------------------------------------------------------------
// Last-frame pruning of the final cross layer: only position t-1 of every sequence feeds the
// head (vap_main.py:316-317 takes [-1]), so its query-side work runs on one row per sequence.
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gather_last(const float* __restrict__ X, const int* __restrict__ tvalid,
                                                     float* __restrict__ Xl, int n_seq, int T) {
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= n_seq) return;
    const int lane = threadIdx.x & 31;
    const int t = tvalid[n >> 1];
    float v[8];
    load_row8(X + ((size_t)n * T + (t - 1)) * kD, lane, v);
    store_row8(Xl + (size_t)n * kD, lane, v);
}
void launch_gather_last(const float* X, const int* tvalid, float* Xl, int n_seq, int T, cudaStream_t st) {
    launch_k(k_gather_last, dim3((n_seq + 7) / 8), dim3(256), 0, st, X, tvalid, Xl, n_seq, T);
}

// One warp per (sequence, head); the query is the newest frame, so every valid key j < t is visible.
// Lane = key in BOTH phases (keys lane and lane + 32 of a 64-key group): the K row and then the V row of a lane's keys
// are fetched with all 32 loads in flight, so a group costs two memory round trips instead of one per 8 keys; the
// 64 partial output sums of a lane are then reduced across the warp by recursive halving (62 shuffles), which leaves
// two output dimensions per lane.
__global__ void __launch_bounds__(128) k_attention_last(AttnArgs a) {
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = a.T;
    const int t = a.tvalid[n >> 1];
    const int kvn = a.sibling ? (n ^ 1) : n;
    float q[64];
    const float4* qp = reinterpret_cast<const float4*>(a.Q + (size_t)n * a.ldq + h * 64);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float4 v = __ldg(qp + i);
        q[4 * i] = v.x; q[4 * i + 1] = v.y; q[4 * i + 2] = v.z; q[4 * i + 3] = v.w;
    }
    const float slope = a.slopes[h];
    const float* kbase = a.K + (size_t)kvn * T * a.ldk + h * 64;
    const float* vbase = a.V + (size_t)kvn * T * a.ldv + h * 64;
    float s[4];
#pragma unroll
    for (int g = 0; g < 2; ++g) {                     // scores of keys lane + 32 * {2g, 2g + 1}
        s[2 * g] = -INFINITY;
        s[2 * g + 1] = -INFINITY;
        if (64 * g < t) {
            const int ja = lane + 64 * g, jb = ja + 32;
            const float4* ka = reinterpret_cast<const float4*>(kbase + (size_t)min(ja, t - 1) * a.ldk);
            const float4* kb = reinterpret_cast<const float4*>(kbase + (size_t)min(jb, t - 1) * a.ldk);
            float4 ra[16], rb[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { ra[i] = ka[i]; rb[i] = kb[i]; }
            float acca = 0.f, accb = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                acca = fmaf(q[4 * i], ra[i].x, acca); acca = fmaf(q[4 * i + 1], ra[i].y, acca);
                acca = fmaf(q[4 * i + 2], ra[i].z, acca); acca = fmaf(q[4 * i + 3], ra[i].w, acca);
                accb = fmaf(q[4 * i], rb[i].x, accb); accb = fmaf(q[4 * i + 1], rb[i].y, accb);
                accb = fmaf(q[4 * i + 2], rb[i].z, accb); accb = fmaf(q[4 * i + 3], rb[i].w, accb);
            }
            if (ja < t) s[2 * g] = acca * 0.0625f + slope * (float)ja;
            if (jb < t) s[2 * g + 1] = accb * 0.0625f + slope * (float)jb;
        }
    }
    float mx = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        s[jj] = (lane + 32 * jj < t) ? expf(s[jj] - mx) : 0.f;
        sum += s[jj];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float o[64];
#pragma unroll
    for (int d = 0; d < 64; ++d) o[d] = 0.f;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        if (64 * g < t) {
            const int ja = lane + 64 * g, jb = ja + 32;
            const float4* va = reinterpret_cast<const float4*>(vbase + (size_t)min(ja, t - 1) * a.ldv);
            const float4* vb = reinterpret_cast<const float4*>(vbase + (size_t)min(jb, t - 1) * a.ldv);
            float4 ra[16], rb[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { ra[i] = va[i]; rb[i] = vb[i]; }
            const float pa = s[2 * g] * inv, pb = s[2 * g + 1] * inv;          // 0 for keys >= t
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                o[4 * i] = fmaf(pa, ra[i].x, fmaf(pb, rb[i].x, o[4 * i]));
                o[4 * i + 1] = fmaf(pa, ra[i].y, fmaf(pb, rb[i].y, o[4 * i + 1]));
                o[4 * i + 2] = fmaf(pa, ra[i].z, fmaf(pb, rb[i].z, o[4 * i + 2]));
                o[4 * i + 3] = fmaf(pa, ra[i].w, fmaf(pb, rb[i].w, o[4 * i + 3]));
            }
        }
    }
    // recursive halving over the warp: after the step with lane-bit `bit` a lane keeps the half of its array selected by
    // that bit (plus what its partner sent for it); 64 -> 32 -> 16 -> 8 -> 4 -> 2 values
    int d0 = 0;
#pragma unroll
    for (int step = 0; step < 5; ++step) {
        const int bit = 16 >> step, half = 32 >> step;
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? o[i] : o[i + half];
            const float keep = up ? o[i + half] : o[i];
            o[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
        if (up) d0 += half;
    }
    float* orow = a.O + (size_t)n * a.ldo + h * 64;
    *reinterpret_cast<float2*>(orow + d0) = make_float2(o[0], o[1]);
}
void launch_attention_last(const AttnArgs& a, cudaStream_t st) {
    launch_k(k_attention_last, dim3(a.n_seq), dim3(128), 0, st, a);
}

// -----------------------------------------------------------------------------------------
// vad = sigmoid(va_classifier(ar_channel output at the last valid frame))
// (vap_main.py:292-293, 313-314).  One warp per (stream, channel).
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vad(const float* __restrict__ X, const int* __restrict__ tvalid,
                                             const float* __restrict__ w, const float* __restrict__ b,
                                             float* __restrict__ out_direct, const IoPtrs* __restrict__ io, int B, int T) {
    pdl_trigger();
    pdl_wait();
    float* out = io ? io->out : out_direct;
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= 2 * B) return;
    const int lane = threadIdx.x & 31;
    const int t = tvalid[n >> 1];
    float v[8], ww[8];
    load_row8(X + ((size_t)n * T + (t - 1)) * kD, lane, v);
    load_row8(w, lane, ww);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(v[i], ww[i], s);
    s = warp_sum(s) + b[0];
    if (lane == 0) out[(n >> 1) * 6 + 4 + (n & 1)] = sigmoidf_(s);
}
void launch_vad(const float* X, const int* tvalid, const float* w, const float* b, float* out, const IoPtrs* io, int B, int T,
                cudaStream_t st) {
    launch_k(k_vad, dim3((2 * B + 7) / 8), dim3(256), 0, st, X, tvalid, w, b, out, io, B, T);
}

// -----------------------------------------------------------------------------------------
// Head: Combinator (modules.py:461-464) on the last valid frame, vap_head / bc_head,
// softmax, codebook aggregation and normalisation (vap_main.py:290-317;
// objective.py:93-110, 186-206; vap_bc_main.py:272-277).  One CTA per stream.
// Also advances the stream's frame counter (last kernel of the step).
// -----------------------------------------------------------------------------------------
constexpr int kHeadWarps = 16;

// NQ dot products of 256-wide weight rows with a vector in smem; all weight loads issued up front
template <int NQ>
__device__ __forceinline__ void warp_dot256(const float* __restrict__ W, int o0, int n_rows, const float* sx, int lane,
                                            float (&y)[NQ]) {
    float ww[NQ][8];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int o = min(o0 + q, n_rows - 1);
        load_row8(W + (size_t)o * kD, lane, ww[q]);
    }
    float xx[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        xx[i] = sx[lane * 4 + i];
        xx[4 + i] = sx[128 + lane * 4 + i];
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(ww[q][i], xx[i], s);
        y[q] = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) y[q] += __shfl_xor_sync(0xffffffffu, y[q], o);
    }
}

__global__ void __launch_bounds__(kHeadWarps * 32) k_head(HeadArgs a) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sx[2][kD];
    __shared__ float sy[2][kD];
    __shared__ float sh[kD];
    __shared__ float sl[kD];
    __shared__ float red[5][8];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = a.tvalid[b];
    if (tid < kD) {
        const size_t r0 = a.compact ? (size_t)(2 * b) : ((size_t)(2 * b) * a.T + (t - 1));
        const size_t r1 = a.compact ? (size_t)(2 * b + 1) : ((size_t)(2 * b + 1) * a.T + (t - 1));
        sx[0][tid] = a.X[r0 * kD + tid];
        sx[1][tid] = a.X[r1 * kD + tid];
    }
    __syncthreads();
    // 16 warps x 8 rows x 2 matrices: each warp has its 16 weight rows of a pass in flight together, two passes
    for (int o0 = warp * 8; o0 < kD; o0 += kHeadWarps * 8) {
        float ya[8], yb[8];
        warp_dot256<8>(a.Wa, o0, kD, sx[0], lane, ya);
        warp_dot256<8>(a.Wb, o0, kD, sx[1], lane, yb);
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                sy[0][o0 + q] = ya[q];
                sy[1][o0 + q] = yb[q];
            }
        }
    }
    __syncthreads();
    if (warp < 2) {                         // one LayerNorm + GELU per channel, shared affine
        float v[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[i] = sy[warp][lane * 4 + i];
            v[4 + i] = sy[warp][128 + lane * 4 + i];
        }
        warp_norm8(v, 1.0f / 256.0f, a.lnw, a.lnb, lane);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            sy[warp][lane * 4 + i] = gelu_erf(v[i]);
            sy[warp][128 + lane * 4 + i] = gelu_erf(v[4 + i]);
        }
    }
    __syncthreads();
    if (tid < kD) {
        sh[tid] = sy[0][tid] + sy[1][tid];
        if (a.comb_tap) a.comb_tap[(size_t)b * kD + tid] = sh[tid];
    }
    __syncthreads();
    for (int o0 = warp * 8; o0 < a.n_out; o0 += kHeadWarps * 8) {
        float y[8];
        warp_dot256<8>(a.Wh, o0, a.n_out, sh, lane, y);
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (o0 + q < a.n_out) sl[o0 + q] = y[q] + a.bh[o0 + q];
        }
    }
    __syncthreads();
    if (a.logits_tap && tid < a.n_out) a.logits_tap[(size_t)b * kD + tid] = sl[tid];
    float* out = (a.io ? a.io->out : a.out) + (size_t)b * 6;
    if (a.head_kind == 0) {
        // softmax over 256 classes, then p[s] = sum_c pi_c * (#active bins of speaker s in the range)
        if (warp < 8) {
            const float lg = sl[tid];
            const float mx = warp_max(lg);
            if (lane == 0) red[0][warp] = mx;
        }
        __syncthreads();
        float mx = red[0][0];
#pragma unroll
        for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[0][i]);
        __syncthreads();
        if (warp < 8) {
            const float e = expf(sl[tid] - mx);
            const int c = tid;
            float vals[5];
            vals[0] = e;
            vals[1] = e * (float)(((c >> 0) & 1) + ((c >> 1) & 1));     // now,    speaker 0: bins 0,1
            vals[2] = e * (float)(((c >> 4) & 1) + ((c >> 5) & 1));     // now,    speaker 1
            vals[3] = e * (float)(((c >> 2) & 1) + ((c >> 3) & 1));     // future, speaker 0: bins 2,3
            vals[4] = e * (float)(((c >> 6) & 1) + ((c >> 7) & 1));     // future, speaker 1
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const float s = warp_sum(vals[k]);
                if (lane == 0) red[k][warp] = s;
            }
        }
        __syncthreads();
        if (tid == 0) {
            float tot[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) s += red[k][i];
                tot[k] = s;
            }
            const float inv = 1.0f / tot[0];
            const float n0 = tot[1] * inv, n1 = tot[2] * inv, f0 = tot[3] * inv, f1 = tot[4] * inv;
            const float dn = n0 + n1 + kEps, df = f0 + f1 + kEps;         // objective.py:205
            out[0] = n0 / dn;
            out[1] = n1 / dn;
            out[2] = f0 / df;
            out[3] = f1 / df;
            // out[4], out[5] (vad) were written by k_vad after the ar_channel layer
        }
    } else {
        if (tid == 0) {
            const float l0 = sl[0], l1 = sl[1], l2 = sl[2];
            const float mx = fmaxf(l0, fmaxf(l1, l2));
            const float e0 = expf(l0 - mx), e1 = expf(l1 - mx), e2 = expf(l2 - mx);
            const float inv = 1.0f / (e0 + e1 + e2);
            out[0] = e1 * inv;          // p_bc_react = softmax[..., 1]   vap_bc_main.py:276
            out[1] = e2 * inv;          // p_bc_emo   = softmax[..., 2]   vap_bc_main.py:277
            out[2] = 0.f; out[3] = 0.f; out[4] = 0.f; out[5] = 0.f;
        }
    }
    if (tid == 0) a.count[a.ids[b]] += 1;
}
void launch_head(const HeadArgs& a, cudaStream_t st) { launch_k(k_head, dim3(a.B), dim3(kHeadWarps * 32), 0, st, a); }

// -----------------------------------------------------------------------------------------
// k_tail: everything behind the window-wide K / V projections of the pruned last cross layer, for the newest frame
// only (modules.py:257-300 restricted to the last position, exact because nothing downstream reads the others:
// vap_main.py:316-317), then Combinator + head + aggregation (modules.py:461-464, vap_main.py:290-317,
// objective.py:186-206) and the frame counter.  Replaces nine dependent launches (2 x LN+Q GEMM, 2 x attention,
// 2 x proj, FFN1, FFN2, head) whose ~10 us each were launch / pipeline-fill latency: the math is 128 rows.
// One CTA owns kTailStreams streams (2 rows each) end to end in true fp32; activations stay in shared memory, the
// 3.4 MB of weights stream through once per CTA as k-major float4 rows, 16 loads in flight per thread.
// -----------------------------------------------------------------------------------------
constexpr int kTailStreams = 2, kTailRows = 2 * kTailStreams;
constexpr int kTailCluster = 4;          // CTAs per row group: each computes a quarter of the output columns of every layer

__device__ __forceinline__ uint32_t tail_crank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void tail_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store one float to the same shared-memory location of CTA `rank` of the cluster (distributed shared memory)
__device__ __forceinline__ void tail_dsm_store(float* local, uint32_t rank, float v) {
    uint32_t la = (uint32_t)__cvta_generic_to_shared(local), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}

// out[r][0..255] = sum_k in[r][k] * WT[k][0..255]   (WT row stride ldw), computed by the FOUR CTAs of the cluster: CTA c
// owns output columns [64c, 64c + 64).  Thread (kq = tid / 16, nq = tid % 16) owns four consecutive outputs over 1/16 of
// K with all its weight loads in flight at once; the 16 partial sums meet in shared memory, the reduced slice is written
// into every CTA's `out` through distributed shared memory, one cluster barrier publishes it.  A single SM pulls only
// ~40 B/clk out of L2: splitting the columns over four SMs is what makes the weight stream short.
__device__ __forceinline__ void tail_lin256(const float* __restrict__ WT, int ldw, int K, const float* in, int ldin, float* part, float* out,
                                            int tid, uint32_t crank, const float* __restrict__ nextWT = nullptr, int next_ldw = 0, int nextK = 0) {
    const int nq = tid & 15, kq = tid >> 4;
    const int kper = K >> 4, k0 = kq * kper;              // 16 (K = 256) or 48 (K = 768)
    if (nextWT) {
        // the weights of the NEXT layer were last touched a step ago and have usually been evicted from L2 by the
        // activation planes of the step: start pulling this thread's rows of them into L2 now
        const int nkper = nextK >> 4;
        const float* np = nextWT + (size_t)(kq * nkper) * next_ldw + 64 * crank + 4 * nq;
        for (int k = 0; k < nkper; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(np + (size_t)k * next_ldw));
    }
    float acc[kTailRows][4];
#pragma unroll
    for (int r = 0; r < kTailRows; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; acc[r][3] = 0.f; }
    const float* wp = WT + (size_t)k0 * ldw + 64 * crank + 4 * nq;
#pragma unroll 4
    for (int k = 0; k < kper; k += 4) {
        float4 w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(k + i) * ldw));
#pragma unroll
        for (int r = 0; r < kTailRows; ++r) {
            const float4 x = *reinterpret_cast<const float4*>(in + r * ldin + k0 + k);
            acc[r][0] = fmaf(x.x, w[0].x, acc[r][0]); acc[r][1] = fmaf(x.x, w[0].y, acc[r][1]); acc[r][2] = fmaf(x.x, w[0].z, acc[r][2]); acc[r][3] = fmaf(x.x, w[0].w, acc[r][3]);
            acc[r][0] = fmaf(x.y, w[1].x, acc[r][0]); acc[r][1] = fmaf(x.y, w[1].y, acc[r][1]); acc[r][2] = fmaf(x.y, w[1].z, acc[r][2]); acc[r][3] = fmaf(x.y, w[1].w, acc[r][3]);
            acc[r][0] = fmaf(x.z, w[2].x, acc[r][0]); acc[r][1] = fmaf(x.z, w[2].y, acc[r][1]); acc[r][2] = fmaf(x.z, w[2].z, acc[r][2]); acc[r][3] = fmaf(x.z, w[2].w, acc[r][3]);
            acc[r][0] = fmaf(x.w, w[3].x, acc[r][0]); acc[r][1] = fmaf(x.w, w[3].y, acc[r][1]); acc[r][2] = fmaf(x.w, w[3].z, acc[r][2]); acc[r][3] = fmaf(x.w, w[3].w, acc[r][3]);
        }
    }
#pragma unroll
    for (int r = 0; r < kTailRows; ++r)
        *reinterpret_cast<float4*>(part + ((kq * kTailRows + r) * 64 + 4 * nq)) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    __syncthreads();
    {   // thread -> (row r = tid / 64, column cidx = tid % 64) of this CTA's slice
        const int r = tid >> 6, cidx = tid & 63;
        float sum = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) sum += part[(q * kTailRows + r) * 64 + cidx];
        float* dst = out + r * kD + 64 * crank + cidx;
#pragma unroll
        for (uint32_t rk = 0; rk < kTailCluster; ++rk) tail_dsm_store(dst, rk, sum);
    }
    tail_cluster_sync();
}

// LayerNorm(256) of the rows in `x` (shared memory) -> `z`; one warp per row
__device__ __forceinline__ void tail_ln_rows(const float* x, float* z, const float* w, const float* b, int warp, int lane, bool gelu) {
    if (warp < kTailRows) {
        float v[8];
        load_row8(x + warp * kD, lane, v);
        warp_norm8(v, 1.0f / 256.0f, w, b, lane);
        if (gelu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
        }
        store_row8(z + warp * kD, lane, v);
    }
    __syncthreads();
}

// newest-frame attention of one (row, head): every valid key j < t is visible.  Lane = key (see k_attention_last).
__device__ __forceinline__ void tail_attend(const float* qs, const float* kbase, const float* vbase, int t, float slope, float* os, int lane) {
    float q[64];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(qs + 4 * i);
        q[4 * i] = v.x; q[4 * i + 1] = v.y; q[4 * i + 2] = v.z; q[4 * i + 3] = v.w;
    }
    float s[4];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        s[2 * g] = -INFINITY;
        s[2 * g + 1] = -INFINITY;
        if (64 * g < t) {
            const int ja = lane + 64 * g, jb = ja + 32;
            const float4* ka = reinterpret_cast<const float4*>(kbase + (size_t)min(ja, t - 1) * 512);
            const float4* kb = reinterpret_cast<const float4*>(kbase + (size_t)min(jb, t - 1) * 512);
            float4 ra[16], rb[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { ra[i] = __ldcg(ka + i); rb[i] = __ldcg(kb + i); }
            float acca = 0.f, accb = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                acca = fmaf(q[4 * i], ra[i].x, acca); acca = fmaf(q[4 * i + 1], ra[i].y, acca);
                acca = fmaf(q[4 * i + 2], ra[i].z, acca); acca = fmaf(q[4 * i + 3], ra[i].w, acca);
                accb = fmaf(q[4 * i], rb[i].x, accb); accb = fmaf(q[4 * i + 1], rb[i].y, accb);
                accb = fmaf(q[4 * i + 2], rb[i].z, accb); accb = fmaf(q[4 * i + 3], rb[i].w, accb);
            }
            if (ja < t) s[2 * g] = acca * 0.0625f + slope * (float)ja;
            if (jb < t) s[2 * g + 1] = accb * 0.0625f + slope * (float)jb;
        }
    }
    float mx = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        s[jj] = (lane + 32 * jj < t) ? expf(s[jj] - mx) : 0.f;
        sum += s[jj];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float o[64];
#pragma unroll
    for (int d = 0; d < 64; ++d) o[d] = 0.f;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        if (64 * g < t) {
            const int ja = lane + 64 * g, jb = ja + 32;
            const float4* va = reinterpret_cast<const float4*>(vbase + (size_t)min(ja, t - 1) * 512);
            const float4* vb = reinterpret_cast<const float4*>(vbase + (size_t)min(jb, t - 1) * 512);
            float4 ra[16], rb[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { ra[i] = __ldcg(va + i); rb[i] = __ldcg(vb + i); }
            const float pa = s[2 * g] * inv, pb = s[2 * g + 1] * inv;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                o[4 * i] = fmaf(pa, ra[i].x, fmaf(pb, rb[i].x, o[4 * i]));
                o[4 * i + 1] = fmaf(pa, ra[i].y, fmaf(pb, rb[i].y, o[4 * i + 1]));
                o[4 * i + 2] = fmaf(pa, ra[i].z, fmaf(pb, rb[i].z, o[4 * i + 2]));
                o[4 * i + 3] = fmaf(pa, ra[i].w, fmaf(pb, rb[i].w, o[4 * i + 3]));
            }
        }
    }
    int d0 = 0;
#pragma unroll
    for (int step = 0; step < 5; ++step) {
        const int bit = 16 >> step, half = 32 >> step;
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? o[i] : o[i + half];
            const float keep = up ? o[i + half] : o[i];
            o[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
        if (up) d0 += half;
    }
    *reinterpret_cast<float2*>(os + d0) = make_float2(o[0], o[1]);
}

// the 16 (row, head) attentions of a row group: CTA c of the cluster takes row c (one warp per head), writes the 256
// outputs to a local scratch row and broadcasts them into every CTA's `ov` row
__device__ __forceinline__ void tail_attention(const float* KV, int sibling, const float* slopes, const float* qv, float* ov, float* scratch, int b0,
                                               int B, int T, const int* tvalid, int warp, int lane, uint32_t crank) {
    const int rr = (int)crank, bs = b0 + (rr >> 1);
    if (warp < kHeads) {
        const int h = warp;
        float* os = scratch + h * 64;
        if (bs < B) {
            const int n = 2 * bs + ((rr & 1) ^ sibling);
            const float* kv = KV + (size_t)n * T * 512 + h * 64;
            tail_attend(qv + rr * kD + h * 64, kv, kv + kD, tvalid[bs], slopes[h], os, lane);
        } else {
            os[lane] = 0.f;
            os[32 + lane] = 0.f;
        }
    }
    __syncthreads();
    {
        const int tid = warp * 32 + lane;
        const float v = scratch[tid];
#pragma unroll
        for (uint32_t rk = 0; rk < kTailCluster; ++rk) tail_dsm_store(ov + rr * kD + tid, rk, v);
    }
    tail_cluster_sync();
}

__global__ void __cluster_dims__(kTailCluster, 1, 1) __launch_bounds__(256) k_tail(TailArgs a) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float tsm[];
    // local buffers
    float* x = tsm;                          // [R][256] residual rows
    float* z = x + kTailRows * kD;           // [R][256] LayerNorm output / scratch
    float* hd = z + kTailRows * kD;          // [R][768] FFN hidden (GELU applied)
    float* part = hd + kTailRows * kFF;      // [16][R][64] split-K partial sums of this CTA's column slice
    // buffers written by every CTA of the cluster (distributed shared memory).  A buffer is written again at the earliest
    // two cluster barriers after its last readers started, so no CTA can still be reading what a faster peer overwrites.
    float* bA = part + 16 * kTailRows * 64;  // [R][256] each
    float* bB = bA + kTailRows * kD;
    float* bC = bB + kTailRows * kD;
    float* bD = bC + kTailRows * kD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t crank = tail_crank();
    const int b0 = (blockIdx.x / kTailCluster) * kTailStreams;
    const int T = a.T;
    // rows: rr = 2 * sl + ch  <->  sequence n = 2 * (b0 + sl) + ch; missing streams of the last cluster run on zero rows
    for (int i = tid; i < kTailRows * kD; i += 256) {
        const int rr = i >> 8, bs = b0 + (rr >> 1);
        x[i] = bs < a.B ? a.Xl[(size_t)(2 * bs + (rr & 1)) * kD + (i & 255)] : 0.f;
    }
    {   // first layer's weight rows of this thread -> L2 (see tail_lin256)
        const float* np = a.WqT + (size_t)((tid >> 4) * 16) * kD + 64 * crank + 4 * (tid & 15);
        for (int k = 0; k < 16; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(np + (size_t)k * kD));
    }
    __syncthreads();
    tail_cluster_sync();                     // every CTA of the cluster is running before the first remote store
    // ---- self attention of the newest frame (modules.py:268-272)
    tail_ln_rows(x, z, a.ln_sa_w, a.ln_sa_b, warp, lane, false);
    tail_lin256(a.WqT, kD, kD, z, kD, part, bA, tid, crank, a.WprojT, kD, kD);
    tail_attention(a.KVs, 0, a.slopes_s, bA, bB, z, b0, a.B, T, a.tvalid, warp, lane, crank);
    tail_lin256(a.WprojT, kD, kD, bB, kD, part, bC, tid, crank, a.WqcT, kD, kD);
    for (int i = tid; i < kTailRows * kD; i += 256) x[i] += bC[i];
    __syncthreads();
    // ---- cross attention: query from LN_src(x), keys / values of the sibling channel (modules.py:276-283)
    tail_ln_rows(x, z, a.ln_src_w, a.ln_src_b, warp, lane, false);
    tail_lin256(a.WqcT, kD, kD, z, kD, part, bA, tid, crank, a.WprojcT, kD, kD);
    tail_attention(a.KVc, 1, a.slopes_c, bA, bB, z, b0, a.B, T, a.tvalid, warp, lane, crank);
    tail_lin256(a.WprojcT, kD, kD, bB, kD, part, bD, tid, crank, a.W1T, kFF, kD);
    for (int i = tid; i < kTailRows * kD; i += 256) x[i] += bD[i];
    __syncthreads();
    // ---- feed forward (modules.py:9-21, 285)
    tail_ln_rows(x, z, a.ln_ff_w, a.ln_ff_b, warp, lane, false);
    for (int j = 0; j < 3; ++j) {
        float* o = (j == 1) ? bC : bA;
        tail_lin256(a.W1T + 256 * j, kFF, kD, z, kD, part, o, tid, crank, j < 2 ? a.W1T + 256 * (j + 1) : a.W2T, j < 2 ? kFF : kD, j < 2 ? kD : kFF);
        for (int i = tid; i < kTailRows * kD; i += 256) hd[(i >> 8) * kFF + 256 * j + (i & 255)] = gelu_erf(o[i]);
    }
    __syncthreads();
    tail_lin256(a.W2T, kD, kFF, hd, kFF, part, bC, tid, crank, a.WaT, kD, kD);
    for (int i = tid; i < kTailRows * kD; i += 256) x[i] += bC[i];
    __syncthreads();
    // ---- Combinator: GELU(LN(h0_a x_ch0)) + GELU(LN(h0_b x_ch1)), one shared LayerNorm (modules.py:461-464)
    tail_lin256(a.WaT, kD, kD, x, kD, part, bA, tid, crank, a.WbT, kD, kD);          // every row through h0_a: the channel-0 rows are used
    tail_lin256(a.WbT, kD, kD, x, kD, part, bD, tid, crank, a.head_kind == 0 ? a.WhT : nullptr, kD, kD);          // every row through h0_b: the channel-1 rows are used
    for (int i = tid; i < kTailRows * kD; i += 256) {
        const int rr = i >> 8;
        z[i] = (rr & 1) ? bD[i] : bA[i];
    }
    __syncthreads();
    tail_ln_rows(z, z, a.comb_lnw, a.comb_lnb, warp, lane, true);
    for (int i = tid; i < kTailRows * kD; i += 256) {          // comb of stream sl -> row sl of hd (ld 256); the other rows zero
        const int rr = i >> 8, cidx = i & 255;
        hd[i] = rr < kTailStreams ? z[(2 * rr) * kD + cidx] + z[(2 * rr + 1) * kD + cidx] : 0.f;
    }
    __syncthreads();
    float* out_base = a.io ? a.io->out : a.out;
    // ---- head (vap_main.py:290-317 ; vap_bc_main.py:272-277)
    if (a.head_kind == 0) {
        tail_lin256(a.WhT, kD, kD, hd, kD, part, bC, tid, crank);     // logits of stream sl in row sl of bC
        if (crank == 0 && warp < kTailStreams && b0 + warp < a.B) {
            float lg[8];
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                lg[i] = bC[warp * kD + lane + 32 * i] + a.bh[lane + 32 * i];
                mx = fmaxf(mx, lg[i]);
            }
            mx = warp_max(mx);
            float vals[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int cc = lane + 32 * i;
                const float e = expf(lg[i] - mx);
                vals[0] += e;
                vals[1] += e * (float)(((cc >> 0) & 1) + ((cc >> 1) & 1));     // now,    speaker 0: bins 0,1
                vals[2] += e * (float)(((cc >> 4) & 1) + ((cc >> 5) & 1));     // now,    speaker 1
                vals[3] += e * (float)(((cc >> 2) & 1) + ((cc >> 3) & 1));     // future, speaker 0: bins 2,3
                vals[4] += e * (float)(((cc >> 6) & 1) + ((cc >> 7) & 1));     // future, speaker 1
            }
#pragma unroll
            for (int k = 0; k < 5; ++k) vals[k] = warp_sum(vals[k]);
            if (lane == 0) {
                float* out = out_base + (size_t)(b0 + warp) * 6;
                const float inv = 1.0f / vals[0];
                const float n0 = vals[1] * inv, n1 = vals[2] * inv, f0 = vals[3] * inv, f1 = vals[4] * inv;
                const float dn = n0 + n1 + kEps, df = f0 + f1 + kEps;         // objective.py:205
                out[0] = n0 / dn;
                out[1] = n1 / dn;
                out[2] = f0 / df;
                out[3] = f1 / df;            // out[4], out[5] (vad) were written after the ar_channel layer
                a.count[a.ids[b0 + warp]] += 1;
            }
        }
    } else {
        if (crank == 0 && warp < kTailStreams && b0 + warp < a.B) {
            float l3[3];
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                float sacc = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) sacc = fmaf(hd[warp * kD + lane + 32 * i], a.WhT[o * kD + lane + 32 * i], sacc);
                l3[o] = warp_sum(sacc) + a.bh[o];
            }
            if (lane == 0) {
                float* out = out_base + (size_t)(b0 + warp) * 6;
                const float mx = fmaxf(l3[0], fmaxf(l3[1], l3[2]));
                const float e0 = expf(l3[0] - mx), e1 = expf(l3[1] - mx), e2 = expf(l3[2] - mx);
                const float inv = 1.0f / (e0 + e1 + e2);
                out[0] = e1 * inv;          // p_bc_react = softmax[..., 1]   vap_bc_main.py:276
                out[1] = e2 * inv;          // p_bc_emo   = softmax[..., 2]   vap_bc_main.py:277
                out[2] = 0.f; out[3] = 0.f; out[4] = 0.f; out[5] = 0.f;
                a.count[a.ids[b0 + warp]] += 1;
            }
        }
    }
    tail_cluster_sync();                     // no CTA exits while a peer may still store into its shared memory
}

constexpr size_t kTailSmem = (size_t)(2 * kTailRows * kD + kTailRows * kFF + 16 * kTailRows * 64 + 4 * kTailRows * kD) * sizeof(float);
void launch_tail(const TailArgs& a, cudaStream_t st) {
    static OncePerDevice once;
    if (once.first()) cudaFuncSetAttribute(k_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTailSmem);
    launch_k(k_tail, dim3(kTailCluster * ((a.B + kTailStreams - 1) / kTailStreams)), dim3(256), kTailSmem, st, a);
}

// -----------------------------------------------------------------------------------------
// Bulk offline scoring helpers (the reference replays a file frame by frame: vap_offline.py:51-73).
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_make_chunks(const float* __restrict__ audio, long long n_samples, int shift, int S, long long first,
                                                     int n_chunks, float* __restrict__ dst) {
    const int q = blockIdx.x;                       // chunk row = ch * n_chunks + b
    const int ch = q / n_chunks, b = q - ch * n_chunks;
    const float* src = audio + (size_t)ch * n_samples + (size_t)shift * (first + b);
    float* d = dst + (size_t)q * S;
    for (int i = threadIdx.x; i < S; i += blockDim.x) d[i] = src[i];
}
void launch_make_chunks(const float* audio, long long n_samples, int shift, int S, long long first, int n_chunks, float* dst, cudaStream_t st) {
    launch_k(k_make_chunks, dim3(2 * n_chunks), dim3(256), 0, st, audio, n_samples, shift, S, first, n_chunks, dst);
}

__global__ void __launch_bounds__(64) k_gather_windows(const float* __restrict__ E, long long n_frames, long long first, int B, int T,
                                                       float* __restrict__ X, int* __restrict__ tvalid, int* __restrict__ ids_out) {
    const int n = blockIdx.x;                       // sequence 2b + ch
    const int b = n >> 1, ch = n & 1;
    const long long f = first + b;
    const int t = (int)((f + 1 < (long long)T) ? f + 1 : (long long)T);
    const float4* src = reinterpret_cast<const float4*>(E + ((size_t)ch * n_frames + (size_t)(f - t + 1)) * kD);
    float4* dst = reinterpret_cast<float4*>(X + (size_t)n * T * kD);
    for (int i = threadIdx.x; i < T * 64; i += blockDim.x) dst[i] = (i < t * 64) ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0 && ch == 0) {
        tvalid[b] = t;
        ids_out[b] = b;
    }
}
void launch_gather_windows(const float* E, long long n_frames, long long first, int B, int T, float* X, int* tvalid, int* ids_out, cudaStream_t st) {
    launch_k(k_gather_windows, dim3(2 * B), dim3(64), 0, st, E, n_frames, first, B, T, X, tvalid, ids_out);
}

// -----------------------------------------------------------------------------------------
// L2 warm-up of the transformer / tail weights, launched on a side branch of the step graph next to the encoder: a
// real-time stream leaves 50 ms between steps, so the weights of the later kernels start a step in HBM, and the per-stream
// cluster kernel otherwise pays the DRAM latency of every weight tile once per launch on its critical path.
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_l2_prefetch(const void* const* __restrict__ ptrs, const unsigned long long* __restrict__ bytes, int n) {
    unsigned long long line = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = 0; i < n; ++i) {
        const unsigned long long nl = (bytes[i] + 127ull) >> 7;
        if (line < nl) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(ptrs[i]) + (line << 7)));
            return;
        }
        line -= nl;
    }
}
void launch_l2_prefetch(const void* const* ptrs, const unsigned long long* bytes, int n, unsigned long long total_lines, cudaStream_t st) {
    launch_k(k_l2_prefetch, dim3((unsigned)((total_lines + 255) / 256)), dim3(256), 0, st, ptrs, bytes, n);
}

}  // namespace vapb
