// Per-stream persistent transformer kernel -- see fused_tf.cuh for the scheme.
//
// CTA anatomy (320 threads, 1 CTA per SM, cluster of 2 CTAs per stream):
//   warps 0-7  workers: A producers (fp32 global -> LayerNorm -> smem transpose -> thread-per-row bf16 hi/lo ->
//              tcgen05.st: the A operand lives in TENSOR MEMORY, so the MMAs read only W from shared memory and
//              shared memory is left for an 8-deep W ring), GEMM epilogues (tcgen05.ld -> GELU / residual ->
//              global), attention, gathers, vad
//   warp 8     TMA producer of the W planes; runs one op AHEAD of the others (weights are constants),
//              so the ring is already full when the cluster barrier in front of a GEMM opens
//   warp 9     TMEM owner + the single thread that issues tcgen05.mma
// All pipelines (W ring, A generations, accumulator slots) carry their phase across ops; the only
// synchronisation between ops is one cluster barrier (release / acquire, covers global memory).
// Activations written earlier in the same launch are read with ld.global.cg (L2), never .nc.
#include "fused_tf.cuh"
#include "tc_ptx.cuh"

#include <cstdlib>
#include <string>

namespace vapb {

namespace {

using namespace tcp;

constexpr int kWorkers = 8;
constexpr int kThreadsF = (kWorkers + 4) * 32;        // 8 workers + TMA + MMA + 2 spare warps (three full warpgroups)
#ifndef VAPB_F_WSTAGES
#define VAPB_F_WSTAGES 6
#endif
constexpr int kWStages = VAPB_F_WSTAGES;
constexpr int kWTile = 64 * kBK * 2;                  // one plane of a 64 (n) x 64 (k) W tile: 8 KB
constexpr int kEpiPitch = 36;
constexpr int kStgPitch = 68;                         // floats per row of the A transpose staging (32 rows x 64 k per quadrant)
constexpr int kStgFloats = 32 * kStgPitch;
// union region: A transpose staging (4 quadrants x 2 buffers), epilogue transpose (8 warps), attention K / V / Q / P
constexpr int kUniBytes = 128 * 1024;              // = two heads x (K hi/lo + V hi/lo) x 16 KB for the tensor-core attention
static_assert(8 * 128 * 128 == kUniBytes && 4 * 32 * kStgPitch * 4 <= 4 * 128 * 128 && 4 * 2 * kStgFloats * 4 <= kUniBytes && kWorkers * 32 * kEpiPitch * 4 <= kUniBytes, "union region too small");
// TMEM columns: A hi plane [0,128), A lo plane [128,256) (bf16x2 per column, K = 256), accumulators [256,512)
constexpr uint32_t kTmALo = 128, kTmAcc = 256;
constexpr int kAccSlots = 4;                          // 4 x 64 accumulator columns
constexpr int kSmemF = kWStages * 2 * kWTile + kUniBytes + 256 /*barriers*/ + 256 /*two op slots*/ + 1024;
static_assert(40 + 16 * kWStages + 16 * kAccSlots + 8 + 48 <= 256, "barrier block too small");

__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void pair_sync(int q) { asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory"); }   // the two warps of a TMEM lane quadrant

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// bounded mbarrier wait that names what it was waiting for (a mis-programmed pipeline must fail loudly, never hang).
// The failure path is a noreturn, non-inlined call: nothing is live across it, so the hot loops keep their registers
// (with the printf inlined, ptxas parked loop state in local memory, and every cluster barrier invalidates L1).
__device__ __noinline__ __attribute__((noreturn)) void fwait_fail(int tag, int oi, uint32_t parity) {
    if ((threadIdx.x & 31) == 0)
        printf("vapb stream kernel: wait %d (1 a_empty 2 acc_full 3 w_empty 4 acc_empty 5 a_full 6 w_full 7 s_full 8 o_full 9 att_in 10 p_ready) timed out, op %d block %d warp %d parity %u\n",
               tag, oi, blockIdx.x, threadIdx.x >> 5, parity);
    __trap();
    for (;;) {}
}
// The wait loop counts polls instead of reading the clock: CS2R shares the XU pipe with the bf16 conversions of the
// worker warps (ncu: XU pipe 61 % busy, 5 % of all executed instructions were clock reads of spinning warps).
__device__ __forceinline__ void fwait(uint32_t bar, uint32_t parity, int tag, int oi) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (VAPB_WAIT_POLLS != 0u && ++polls > VAPB_WAIT_POLLS) fwait_fail(tag, oi, parity);      // each failed try_wait suspends the thread for a while: seconds
    }
}

struct Ctx {
    uint32_t wbase, bars, tmem_base;
    float* uni;            // union region (generic pointer)
    FOpFields* opslot;     // two shared-memory slots: fields of the current / next op
    int tid, warp, lane;
    int b, r;              // stream index in the batch, CTA rank in the cluster
    int m0, rows;          // first global row and valid rows of this CTA's M tile
    int T, t, mode, oi;
    __device__ __forceinline__ uint32_t w_hi(int s) const { return wbase + (uint32_t)s * 2 * kWTile; }
    __device__ __forceinline__ uint32_t w_lo(int s) const { return w_hi(s) + kWTile; }
    __device__ __forceinline__ uint32_t a_full(int kb) const { return bars + 8u * kb; }
    __device__ __forceinline__ uint32_t a_empty() const { return bars + 32u; }
    __device__ __forceinline__ uint32_t w_full(int s) const { return bars + 40u + 8u * s; }
    __device__ __forceinline__ uint32_t w_empty(int s) const { return bars + 40u + 8u * (kWStages + s); }
    __device__ __forceinline__ uint32_t acc_full(int i) const { return bars + 40u + 16u * kWStages + 8u * i; }
    __device__ __forceinline__ uint32_t acc_empty(int i) const { return bars + 40u + 16u * kWStages + 8u * (kAccSlots + i); }
    __device__ __forceinline__ uint32_t tmem_slot() const { return bars + 40u + 16u * kWStages + 16u * kAccSlots; }
    // tensor-core attention: inputs staged / S complete / P written (per head) / O complete (per head)
    __device__ __forceinline__ uint32_t att_in() const { return tmem_slot() + 8u; }
    __device__ __forceinline__ uint32_t s_full() const { return tmem_slot() + 16u; }
    __device__ __forceinline__ uint32_t p_ready(int x) const { return tmem_slot() + 24u + 8u * x; }
    __device__ __forceinline__ uint32_t o_full(int x) const { return tmem_slot() + 40u + 8u * x; }
    __device__ __forceinline__ uint32_t uni_u32() const { return wbase + kWStages * 2 * kWTile; }
};

// exact-erf GELU with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, the level of erff itself): one fast
// reciprocal, one ex2 and five FMAs instead of erff's ~22 instructions; 192 GELUs per thread in an FFN1 epilogue
__device__ __forceinline__ float gelu_as(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
    float pl = fmaf(1.061405429f, t, -1.453152027f);
    pl = fmaf(pl, t, 1.421413741f);
    pl = fmaf(pl, t, -0.284496736f);
    pl = fmaf(pl, t, 0.254829592f);
    const float er = 1.0f - pl * t * __expf(-z * z);          // erf(|x| / sqrt 2)
    return 0.5f * x + 0.5f * fabsf(x) * er;                   // 0.5 x (1 + sign(x) erf(|x| / sqrt 2))
}

// One 32 x 32 block of the accumulator: TMEM -> registers -> per-warp smem transpose -> (GELU, + R) -> global.
template <bool HAS_R, bool GELU>
__device__ __forceinline__ void epi_block(const uint32_t (&raw)[32], float* tbuf, int lane, int rows_valid, const float* Rb,
                                          float* Cb, int ld, int ncol) {
    const int sub = lane >> 3;
    const int c4 = (lane & 7) * 4;
#pragma unroll
    for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(tbuf + lane * kEpiPitch + 4 * q) =
            make_float4(__uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]), __uint_as_float(raw[4 * q + 2]),
                        __uint_as_float(raw[4 * q + 3]));
    __syncwarp();
    float4 res[8];
    if (HAS_R) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int rr = 4 * it + sub;
            res[it] = (rr < rows_valid) ? ldcg4(Rb + (size_t)rr * ld + ncol + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int rr = 4 * it + sub;
        float4 x = *reinterpret_cast<const float4*>(tbuf + rr * kEpiPitch + c4);
        if (GELU) {
            x.x = gelu_as(x.x); x.y = gelu_as(x.y); x.z = gelu_as(x.z); x.w = gelu_as(x.w);
        }
        if (HAS_R) {
            x.x += res[it].x; x.y += res[it].y; x.z += res[it].z; x.w += res[it].w;
        }
        if (rr < rows_valid) *reinterpret_cast<float4*>(Cb + (size_t)rr * ld + ncol + c4) = x;
    }
    __syncwarp();
}

// One 64-wide k-block of a 128-row operand tile: coalesced-layout registers (warp (q, hf) holds rows 32q + 16hf + 4i + lg,
// 8 floats at column 8 * chunk) -> staging tile of the quadrant -> thread-per-row bf16 hi / lo -> 16 + 16 TMEM columns.
__device__ __forceinline__ void stage_to_tmem(const Ctx& c, const float4 (&v)[4][2], float* sb, uint32_t tm_hi, uint32_t tm_lo) {
    const int q = c.warp & 3, hf = c.warp >> 2, lg = c.lane >> 3, chunk = c.lane & 7;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float* dst = sb + (16 * hf + 4 * i + lg) * kStgPitch + chunk * 8;
        *reinterpret_cast<float4*>(dst) = v[i][0];
        *reinterpret_cast<float4*>(dst + 4) = v[i][1];
    }
    pair_sync(q);
    const float* src = sb + c.lane * kStgPitch + 32 * hf;
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float4 f = *reinterpret_cast<const float4*>(src + 4 * e);
        split2(f.x, f.y, hi[2 * e], lo[2 * e]);
        split2(f.z, f.w, hi[2 * e + 1], lo[2 * e + 1]);
    }
    tmem_st16(tm_hi + (uint32_t)(16 * hf), hi);
    tmem_st16(tm_lo + (uint32_t)(16 * hf), lo);
}

// ---- GEMM, worker side: write the A operand of every 256-wide K chunk to tensor memory, then drain the accumulators ----
// Load layout: warp (q = warp & 3, hf = warp >> 2) holds rows 32q + 16hf + [0,16) of the tile, all 256 columns
// (8 lanes per row, 128-bit coalesced loads; LayerNorm statistics by 8-lane shuffles).  tcgen05.st wants one thread
// per row (TMEM lane = row, reachable only from warps with warp % 4 == q), so every 64-wide k-block is transposed
// through a [32 rows][64] staging tile shared by the two warps of the quadrant: thread l then owns row 32q + l,
// columns 32hf + [0,32) of the k-block, splits them into bf16 hi / lo pairs and stores 16 + 16 packed columns.
template <bool LN>
__device__ __forceinline__ void gemm_workers(const Ctx& c, const FOpFields& op, int n_begin, int ns, int gst, int ga, long long* d2) {
    const int kch = op.K >> 8;
    const int q = c.warp & 3, hf = c.warp >> 2, lg = c.lane >> 3, chunk = c.lane & 7;
    float* stg = c.uni + q * (2 * kStgFloats);
    const uint32_t tm_q = c.tmem_base + ((uint32_t)(q * 32) << 16);
    for (int kc = 0; kc < kch; ++kc) {
        float4 vr[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = 32 * q + 16 * hf + 4 * i + lg;
            if (r < c.rows) {
                const float* p = op.A + (size_t)(c.m0 + r) * op.lda + kc * 256 + chunk * 8;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    vr[j][i][0] = ldcg4(p + j * kBK);
                    vr[j][i][1] = ldcg4(p + j * kBK + 4);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    vr[j][i][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                    vr[j][i][1] = vr[j][i][0];
                }
            }
        }
        if constexpr (LN) {
            // LayerNorm(256) of the 4 rows this thread shares with its 8-lane group: two-pass statistics,
            // deviations kept in place, y = (x - mean) * rstd * w + b
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    sum += (vr[j][i][0].x + vr[j][i][0].y) + (vr[j][i][0].z + vr[j][i][0].w) + (vr[j][i][1].x + vr[j][i][1].y) +
                           (vr[j][i][1].z + vr[j][i][1].w);
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                sum += __shfl_xor_sync(0xffffffffu, sum, 4);
                const float mean = sum * (1.0f / 256.0f);
                float sq0 = 0.f, sq1 = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4& a = vr[j][i][0];
                    float4& b = vr[j][i][1];
                    a.x -= mean; a.y -= mean; a.z -= mean; a.w -= mean;
                    b.x -= mean; b.y -= mean; b.z -= mean; b.w -= mean;
                    sq0 = fmaf(a.x, a.x, sq0); sq1 = fmaf(a.y, a.y, sq1); sq0 = fmaf(a.z, a.z, sq0); sq1 = fmaf(a.w, a.w, sq1);
                    sq0 = fmaf(b.x, b.x, sq0); sq1 = fmaf(b.y, b.y, sq1); sq0 = fmaf(b.z, b.z, sq0); sq1 = fmaf(b.w, b.w, sq1);
                }
                float sq = sq0 + sq1;
                sq += __shfl_xor_sync(0xffffffffu, sq, 1);
                sq += __shfl_xor_sync(0xffffffffu, sq, 2);
                sq += __shfl_xor_sync(0xffffffffu, sq, 4);
                const float rstd = 1.0f / sqrtf(sq * (1.0f / 256.0f) + 1e-5f);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4& a = vr[j][i][0];
                    float4& b = vr[j][i][1];
                    a.x *= rstd; a.y *= rstd; a.z *= rstd; a.w *= rstd;
                    b.x *= rstd; b.y *= rstd; b.z *= rstd; b.w *= rstd;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(op.ln_w + j * kBK + chunk * 8));
                const float4 w1 = __ldg(reinterpret_cast<const float4*>(op.ln_w + j * kBK + chunk * 8 + 4));
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(op.ln_b + j * kBK + chunk * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(op.ln_b + j * kBK + chunk * 8 + 4));
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4& a = vr[j][i][0];
                    float4& b = vr[j][i][1];
                    a.x = fmaf(a.x, w0.x, b0.x); a.y = fmaf(a.y, w0.y, b0.y); a.z = fmaf(a.z, w0.z, b0.z); a.w = fmaf(a.w, w0.w, b0.w);
                    b.x = fmaf(b.x, w1.x, b1.x); b.y = fmaf(b.y, w1.y, b1.y); b.z = fmaf(b.z, w1.z, b1.z); b.w = fmaf(b.w, w1.w, b1.w);
                }
            }
        }
        if (d2 && kc == 0) {
            asm volatile("" ::"f"(vr[3][3][1].w));
            d2[1] = clock64();                 // A loads landed (+ LayerNorm done)
        }
        // the MMAs of the previous A generation must have finished reading the A columns of tensor memory
        fwait(c.a_empty(), (uint32_t)((ga + kc) & 1) ^ 1u, 1, c.oi);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            stage_to_tmem(c, vr[j], stg + (j & 1) * kStgFloats, tm_q + (uint32_t)(j * 32), tm_q + kTmALo + (uint32_t)(j * 32));
            tmem_st_wait();
            tc_fence_before();
            if (c.lane == 0) mbar_arrive(c.a_full(j));
        }
    }
    if (d2) d2[2] = clock64();                 // A planes stored
    // ---- epilogue: one 32 x 32 block per warp and 64-wide subtile
    const int quad = q, half = hf;
    const int rows_valid = min(32, c.rows - quad * 32);
    float* tbuf = c.uni + c.warp * (32 * kEpiPitch);     // aliases the A staging: every warp is past it once an accumulator is complete
    const uint32_t tm_row = c.tmem_base + ((uint32_t)(quad * 32) << 16);
    const size_t rowoff = (size_t)(c.m0 + quad * 32) * op.ldc;
    float* Cb = op.C + rowoff;
    const float* Rb = op.R ? op.R + rowoff : nullptr;
    for (int st = 0; st < ns; ++st) {
        const int g = gst + st, slot = g & (kAccSlots - 1);
        fwait(c.acc_full(slot), (uint32_t)(g / kAccSlots) & 1u, 2, c.oi);
        tc_fence_after();
        if (d2 && st == 0) d2[3] = clock64();          // first accumulator complete
        if (d2 && st == ns - 1) d2[4] = clock64();     // last accumulator complete
        uint32_t raw[32];
        tmem_ld32(tm_row + kTmAcc + (uint32_t)(slot * 64 + half * 32), raw);
        tc_fence_before();
        if (c.lane == 0) mbar_arrive(c.acc_empty(slot));     // slot may be overwritten by a later subtile / op
        const int ncol = n_begin + st * 64 + half * 32;
        if (Rb) {
            if (op.act == 1) epi_block<true, true>(raw, tbuf, c.lane, rows_valid, Rb, Cb, op.ldc, ncol);
            else epi_block<true, false>(raw, tbuf, c.lane, rows_valid, Rb, Cb, op.ldc, ncol);
        } else {
            if (op.act == 1) epi_block<false, true>(raw, tbuf, c.lane, rows_valid, Rb, Cb, op.ldc, ncol);
            else epi_block<false, false>(raw, tbuf, c.lane, rows_valid, Rb, Cb, op.ldc, ncol);
        }
    }
    if (d2) d2[5] = clock64();                 // epilogue of warp 0 done
}

// ---- causal ALiBi attention (modules.py:82-110, 170-212) on the tensor cores, two heads per round ----
//   S = Q K^T (bf16 x3, A = Q in tensor memory, B = K in smem)  ->  thread-per-row masked softmax with the ALiBi
//   bias, P written over S as bf16 hi / lo  ->  O = P V (A = P in tensor memory, B = V in smem, MN-major: V stays
//   [key][dim], no transpose)  ->  O rows to global.
// Tile rows: mode 0 puts sequence 0 in rows 0..63 and sequence 1 in rows 64..127 (T <= 64 valid each), so a TMEM lane
// quadrant never mixes sequences and S is block diagonal (the off-diagonal blocks of P are written as zeros);
// mode 1 has one sequence per CTA (rows 0..T-1).  K / V tile rows use the same mapping (sibling channel when
// cross-attending).  TMEM columns: Q_x hi/lo at 64x (+32), S_x / P_x at 128 + 128x, O_x at 384 + 64x.
// P of the 32-key chunk c lives in the 32 columns of S chunk c (16 hi + 16 lo), so overwriting S in place is safe.
constexpr uint32_t kTmQ = 0, kTmS = 128, kTmO = 384;
constexpr int kAttPlane = 128 * 128;                  // bytes of one bf16 plane of a [128 keys][64 dims] tile
__host__ __device__ constexpr uint32_t make_idesc_bmn(int bn) { return make_idesc(bn) | (1u << 16); }   // B operand MN-major

struct AttMap {            // tile row -> (sequence, position)
    int mode, b, r, T;
    __device__ __forceinline__ int seq(int rr) const { return mode == 0 ? 2 * b + (rr >> 6) : 2 * b + r; }
    __device__ __forceinline__ int pos(int rr) const { return mode == 0 ? (rr & 63) : rr; }
};

__device__ __forceinline__ void att_load_rows(const Ctx& c, const AttMap& am, const float* base, int ld, int h, int sib, int limit,
                                              float4 (&v)[4][2]) {
    const int q = c.warp & 3, hf = c.warp >> 2, lg = c.lane >> 3, chunk = c.lane & 7;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rr = 32 * q + 16 * hf + 4 * i + lg;
        const int pos = am.pos(rr);
        if (pos < limit) {
            const float* p = base + ((size_t)(am.seq(rr) ^ sib) * am.T + pos) * ld + h * 64 + chunk * 8;
            v[i][0] = ldcg4(p);
            v[i][1] = ldcg4(p + 4);
        } else {
            v[i][0] = make_float4(0.f, 0.f, 0.f, 0.f);
            v[i][1] = v[i][0];
        }
    }
}
// rows of a [128][64] operand tile -> bf16 hi / lo planes in the SWIZZLE_128B layout (row = 128 bytes)
__device__ __forceinline__ void att_store_planes(const Ctx& c, const float4 (&v)[4][2], uint32_t hi_plane, uint32_t lo_plane) {
    const int q = c.warp & 3, hf = c.warp >> 2, lg = c.lane >> 3, chunk = c.lane & 7;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = 32 * q + 16 * hf + 4 * i + lg;
        uint32_t h[4], l[4];
        split2(v[i][0].x, v[i][0].y, h[0], l[0]);
        split2(v[i][0].z, v[i][0].w, h[1], l[1]);
        split2(v[i][1].x, v[i][1].y, h[2], l[2]);
        split2(v[i][1].z, v[i][1].w, h[3], l[3]);
        const uint32_t off = (uint32_t)r * 128u + ((uint32_t)(chunk ^ (r & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hi_plane + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3])
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lo_plane + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3])
                     : "memory");
    }
}

__device__ __forceinline__ int att_rounds(const Ctx& c) { return c.mode == 0 ? 1 : 2; }

__device__ __forceinline__ void attention_workers(const Ctx& c, const FOpFields& op, int na, long long* d2) {
    const AttMap am{c.mode, c.b, c.r, c.T};
    const int q = c.warp & 3, x = c.warp >> 2;             // softmax / epilogue: quadrant q of head x
    const uint32_t tm_q = c.tmem_base + ((uint32_t)(q * 32) << 16);
    // smem: K_x hi / lo at x * 2 planes, V_x hi / lo at (4 + 2x) planes; the Q staging tiles (one per quadrant) alias V
    float* stg = c.uni + kAttPlane + q * kStgFloats;       // float offset kAttPlane = byte offset 4 planes
    const uint32_t ub = c.uni_u32();
    for (int round = 0; round < att_rounds(c); ++round) {
        const uint32_t par = (uint32_t)(na + round) & 1u;
        const int hbase = c.mode == 0 ? 2 * c.r : 2 * round;
        // ---- stage Q (tensor memory), K and V (shared memory) of both heads; the loads of the next tensor are in
        //      flight while the previous one is converted
        {
            float4 va[2][4][2], vb[2][4][2];
#pragma unroll
            for (int hx = 0; hx < 2; ++hx) att_load_rows(c, am, op.Q, op.ldq, hbase + hx, 0, c.t, va[hx]);
#pragma unroll
            for (int hx = 0; hx < 2; ++hx) att_load_rows(c, am, op.Kp, op.ldk, hbase + hx, op.sibling, c.t, vb[hx]);
            if (d2) {
                asm volatile("" ::"f"(va[1][3][1].w));
                d2[1] = clock64();             // Q loads landed
            }
#pragma unroll
            for (int hx = 0; hx < 2; ++hx) {
                if (hx) pair_sync(q);          // the partner warp has read head 0 out of the staging tile
                stage_to_tmem(c, va[hx], stg, tm_q + kTmQ + 64u * hx, tm_q + kTmQ + 64u * hx + 32u);
            }
#pragma unroll
            for (int hx = 0; hx < 2; ++hx) att_load_rows(c, am, op.V, op.ldv, hbase + hx, op.sibling, c.t, va[hx]);
#pragma unroll
            for (int hx = 0; hx < 2; ++hx) att_store_planes(c, vb[hx], ub + hx * 2 * kAttPlane, ub + hx * 2 * kAttPlane + kAttPlane);
            workers_sync();                    // every warp is past the Q staging tiles, which alias the V planes
#pragma unroll
            for (int hx = 0; hx < 2; ++hx)
                att_store_planes(c, va[hx], ub + (4 + 2 * hx) * kAttPlane, ub + (5 + 2 * hx) * kAttPlane);
            fence_proxy_async();
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (c.lane == 0) mbar_arrive(c.att_in());
            if (d2) d2[2] = clock64();         // Q / K / V staged
        }
        // ---- softmax of head x, rows of quadrant q: thread = row
        {
            const int rr = 32 * q + c.lane;
            const int i = am.pos(rr);
            const bool rowv = i < c.t;
            const int blk = c.mode == 0 ? (q >> 1) : 0;            // 64-key block that holds this row's sequence
            const int cfirst = 2 * blk;                            // first 32-key chunk of the sequence
            const int nch = (c.mode == 0 ? (q & 1) : q) + 1;       // chunks with visible keys (causal)
            const float slope = __ldg(op.slopes + hbase + x);
            const uint32_t tS = tm_q + kTmS + 128u * x;
            fwait(c.s_full(), par, 7, c.oi);
            tc_fence_after();
            if (d2) d2[3] = clock64();         // S complete
            if (nch <= 2) {
                // at most 64 visible keys (always the case for T <= 64): the scores stay in registers, one pass
                uint32_t r0[32], r1[32];
                tmem_ld32(tS + 32u * cfirst, r0);
                if (nch == 2) tmem_ld32(tS + 32u * (cfirst + 1), r1);
                float m = -INFINITY;
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const float s0 = (rowv && e <= i) ? __uint_as_float(r0[e]) * 0.0625f + slope * (float)e : -INFINITY;
                    const float s1 = (nch == 2 && rowv && 32 + e <= i) ? __uint_as_float(r1[e]) * 0.0625f + slope * (float)(32 + e) : -INFINITY;
                    r0[e] = __float_as_uint(s0);
                    r1[e] = __float_as_uint(s1);
                    m = fmaxf(m, fmaxf(s0, s1));
                }
                const float mm = (m > -INFINITY) ? m : 0.f;
                float l = 0.f;
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const float e0 = __expf(__uint_as_float(r0[e]) - mm);      // exp(-inf) = 0 for masked keys
                    const float e1 = __expf(__uint_as_float(r1[e]) - mm);
                    r0[e] = __float_as_uint(e0);
                    r1[e] = __float_as_uint(e1);
                    l += e0 + e1;
                }
                const float inv = (l > 0.f) ? 1.0f / l : 0.f;
                for (int cc = 0; cc < 4; ++cc) {
                    uint32_t hi[16], lo[16];
                    const int lc = cc - cfirst;
                    if (lc == 0) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) split2(__uint_as_float(r0[2 * e]) * inv, __uint_as_float(r0[2 * e + 1]) * inv, hi[e], lo[e]);
                    } else if (lc == 1 && nch == 2) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) split2(__uint_as_float(r1[2 * e]) * inv, __uint_as_float(r1[2 * e + 1]) * inv, hi[e], lo[e]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e) { hi[e] = 0u; lo[e] = 0u; }
                    }
                    tmem_st16(tS + 32u * cc, hi);
                    tmem_st16(tS + 32u * cc + 16u, lo);
                }
            } else {
            float m = -INFINITY, l = 0.f;
            for (int cc = 0; cc < nch; ++cc) {
                uint32_t raw[32];
                tmem_ld32(tS + 32u * (cfirst + cc), raw);
                float cm = -INFINITY;
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int j = 32 * cc + e;
                    const float sv = (rowv && j <= i) ? __uint_as_float(raw[e]) * 0.0625f + slope * (float)j : -INFINITY;
                    raw[e] = __float_as_uint(sv);
                    cm = fmaxf(cm, sv);
                }
                const float mn = fmaxf(m, cm);
                float acc = 0.f;
                if (mn > -INFINITY) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) acc += __expf(__uint_as_float(raw[e]) - mn);
                    l = l * __expf(m - mn) + acc;
                }
                m = mn;
            }
            const float inv = (l > 0.f) ? 1.0f / l : 0.f;
            const float mm = (m > -INFINITY) ? m : 0.f;
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t hi[16], lo[16];
                const int lc = cc - cfirst;               // chunk index inside the row's own sequence
                if (lc >= 0 && lc < nch) {
                    uint32_t raw[32];
                    tmem_ld32(tS + 32u * cc, raw);
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int j0 = 32 * lc + 2 * e, j1 = j0 + 1;
                        const float p0 = (rowv && j0 <= i) ? __expf(__uint_as_float(raw[2 * e]) * 0.0625f + slope * (float)j0 - mm) * inv : 0.f;
                        const float p1 = (rowv && j1 <= i) ? __expf(__uint_as_float(raw[2 * e + 1]) * 0.0625f + slope * (float)j1 - mm) * inv : 0.f;
                        split2(p0, p1, hi[e], lo[e]);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) { hi[e] = 0u; lo[e] = 0u; }
                }
                tmem_st16(tS + 32u * cc, hi);
                tmem_st16(tS + 32u * cc + 16u, lo);
            }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (c.lane == 0) mbar_arrive(c.p_ready(x));
            if (d2) d2[4] = clock64();         // P written
        }
        // ---- O of head x, rows of quadrant q -> global
        {
            fwait(c.o_full(x), par, 8, c.oi);
            tc_fence_after();
            if (d2) d2[7] = clock64();         // O complete
            const int seq = c.mode == 0 ? 2 * c.b + (q >> 1) : 2 * c.b + c.r;
            const int i0 = c.mode == 0 ? 32 * (q & 1) : 32 * q;
            const int rows_valid = min(32, c.T - i0);
            float* tbuf = c.uni + c.warp * (32 * kEpiPitch);      // aliases K of head 0 (both S products are complete)
            float* Cb = op.O + ((size_t)seq * c.T + i0) * op.ldo;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t raw[32];
                tmem_ld32(tm_q + kTmO + 64u * x + 32u * half, raw);
                epi_block<false, false>(raw, tbuf, c.lane, rows_valid, nullptr, Cb, op.ldo, (hbase + x) * 64 + 32 * half);
            }
            if (d2) d2[5] = clock64();         // O written
        }
        if (round + 1 < att_rounds(c)) {
            tc_fence_before();
            workers_sync();                    // K / V / staging tiles and the TMEM columns are reused by the next round
        }
    }
}

// MMA side of the attention (one elected thread)
__device__ __forceinline__ void attention_mma(const Ctx& c, int na) {
    const uint32_t ub = c.uni_u32();
    constexpr uint32_t idesc_s = make_idesc(128);
    constexpr uint32_t idesc_o = make_idesc_bmn(64);
    const int nk = c.mode == 0 ? 8 : (c.T + 15) / 16;      // 16-key steps of P V
    for (int round = 0; round < att_rounds(c); ++round) {
        const uint32_t par = (uint32_t)(na + round) & 1u;
        fwait(c.att_in(), par, 9, c.oi);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t koff = (uint32_t)k * kUmmaK * 2;
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {
#pragma unroll
                for (int hx = 0; hx < 2; ++hx) {
                    const uint32_t qh = c.tmem_base + kTmQ + 64u * hx + 8u * k, ql = qh + 32u;
                    const uint32_t kh = ub + hx * 2 * kAttPlane + koff, kl = kh + kAttPlane;
                    const uint32_t acc = c.tmem_base + kTmS + 128u * hx;
                    if (prod == 0) umma_bf16_ta(acc, ql, make_desc(kh), idesc_s, k ? 1u : 0u);
                    else if (prod == 1) umma_bf16_ta(acc, qh, make_desc(kl), idesc_s, 1u);
                    else umma_bf16_ta(acc, qh, make_desc(kh), idesc_s, 1u);
                }
            }
        }
        umma_commit(c.s_full());
        fwait(c.p_ready(0), par, 10, c.oi);
        fwait(c.p_ready(1), par, 10, c.oi);
        tc_fence_after();
        for (int kk = 0; kk < nk; ++kk) {
            // keys [16 kk, +16): chunk kk / 2, half kk % 2 -> 8 packed columns of the hi and of the lo half of the chunk;
            // the two heads alternate so that dependent MMAs on one accumulator are not back to back
            const uint32_t voff = (uint32_t)kk * 2048u;        // 16 key rows of 128 bytes
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {
#pragma unroll
                for (int hx = 0; hx < 2; ++hx) {
                    const uint32_t acc = c.tmem_base + kTmO + 64u * hx;
                    const uint32_t vh = ub + (4 + 2 * hx) * kAttPlane + voff, vl = vh + kAttPlane;
                    const uint32_t ph = c.tmem_base + kTmS + 128u * hx + 32u * (kk >> 1) + 8u * (kk & 1), pl = ph + 16u;
                    if (prod == 0) umma_bf16_ta(acc, pl, make_desc(vh), idesc_o, kk ? 1u : 0u);
                    else if (prod == 1) umma_bf16_ta(acc, ph, make_desc(vl), idesc_o, 1u);
                    else umma_bf16_ta(acc, ph, make_desc(vh), idesc_o, 1u);
                }
            }
        }
        umma_commit(c.o_full(0));
        umma_commit(c.o_full(1));
    }
}

// ---- op shape of this CTA (loaded values are shuffled so that the compiler knows they are warp-uniform) ----
struct OpShape { int kind, ns, kch, n_begin; };
__device__ __forceinline__ OpShape op_shape(const Ctx& c, const FOpFields& op) {
    OpShape o;
    o.kind = __shfl_sync(0xffffffffu, op.kind, 0);
    o.ns = 0; o.kch = 0; o.n_begin = 0;
    if (o.kind == FOP_GEMM) {
        o.kch = op.K >> 8;
        if (c.mode == 0) { o.ns = op.N >> 7; o.n_begin = c.r * (op.N >> 1); }
        else { o.ns = op.N >> 6; }
    }
    o.ns = __shfl_sync(0xffffffffu, o.ns, 0);
    o.kch = __shfl_sync(0xffffffffu, o.kch, 0);
    o.n_begin = __shfl_sync(0xffffffffu, o.n_begin, 0);
    return o;
}
__device__ __forceinline__ long long* fine_stamps(const FusedParams& p, int oi) {
    // fine stamps of one op (p.dbg_op) of cluster 0 / CTA 0: slots 40.. of the clock buffer
    return (p.dbg != nullptr && blockIdx.x == 0 && oi == p.dbg_op) ? p.dbg + 40 : nullptr;
}

// =========================== workers (warps 0-7) ===========================
__device__ __forceinline__ void workers_loop(Ctx& c, const FusedParams& p, int id, int cnt) {
    int gst = 0, ga = 0, na = 0;     // running counters: accumulator subtiles, A generations, attention rounds
    const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && c.tid == 0;
    // programmatic dependent launch: set-up and the first weight tiles overlap the tail of the kernel in front; the
    // workers are the only warps that read what it produced (the ring)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int oi = 0; oi < p.n_ops; ++oi) {
        const FOpFields& op = c.opslot[oi & 1];      // copied one op ahead by the MMA warp, visible through the cluster barrier
        c.oi = oi;
        const OpShape sh = op_shape(c, op);
        if (dbg) p.dbg[oi] = clock64();
        long long* d2 = c.tid == 0 ? fine_stamps(p, oi) : nullptr;
        if (d2) d2[0] = clock64();
        if (sh.kind == FOP_GEMM) {
            if (op.ln_w) gemm_workers<true>(c, op, sh.n_begin, sh.ns, gst, ga, d2);
            else gemm_workers<false>(c, op, sh.n_begin, sh.ns, gst, ga, d2);
            gst += sh.ns;
            ga += sh.kch;
        } else if (sh.kind == FOP_ATTN) {
            attention_workers(c, op, na, d2);
            na += att_rounds(c);
        } else if (sh.kind == FOP_GATHER_RING) {
            // X rows of channel r = ring rows oldest first, zero rows above t (vap_main.py:274-283)
            const int ch = c.r;
            const float* rg = p.ring + ((size_t)id * 2 + ch) * p.T * kD;      // (row layout note: a lane owns 8 consecutive floats below)
            float* xo = p.X + (size_t)(2 * c.b + ch) * p.T * kD;
            // thread = (row j0 + 4 u, float4 q): 8 rows per pass with all loads in flight (a plain loop pays one L2 round
            // trip per row)
            const int q4 = c.tid & 63, j0 = c.tid >> 6;
            const int jnew = p.ds_part ? c.t - 1 : -1;           // newest frame: produced below, not yet in the ring
            for (int jb = 0; jb < p.T; jb += 32) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = jb + j0 + 4 * u;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j < c.t && j != jnew) {
                        const int slot = (cnt - c.t + j) % p.T;
                        v[u] = ldcg4(rg + (size_t)slot * kD + 4 * q4);
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = jb + j0 + 4 * u;
                    if (j < p.T && j != jnew) *reinterpret_cast<float4*>(xo + (size_t)j * kD + 4 * q4) = v[u];
                }
            }
            if (p.ds_part && c.warp == kWorkers - 1) {
                // downsample tail of channel r: sum the split-K partials, LayerNorm (biased variance), exact GELU
                const int n = 2 * c.b + ch;
                const float* pr = p.ds_part + (size_t)n * kD + 8 * c.lane;
                // partial z lives ds_stride floats behind partial z - 1; 8 partials (16 loads) in flight per pass
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b4 = a;
                for (int z0 = 0; z0 < p.ds_nsplit; z0 += 8) {
                    float4 pa[8], pb[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float* q8 = pr + (size_t)min(z0 + u, p.ds_nsplit - 1) * p.ds_stride;
                        pa[u] = ldcg4(q8);
                        pb[u] = ldcg4(q8 + 4);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (z0 + u < p.ds_nsplit) {
                            a.x += pa[u].x; a.y += pa[u].y; a.z += pa[u].z; a.w += pa[u].w;
                            b4.x += pb[u].x; b4.y += pb[u].y; b4.z += pb[u].z; b4.w += pb[u].w;
                        }
                    }
                }
                float v8[8] = {a.x, a.y, a.z, a.w, b4.x, b4.y, b4.z, b4.w};
                float sum = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) sum += v8[i];
                const float mean = warp_sum_f(sum) * (1.0f / 256.0f);
                float sq = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float d = v8[i] - mean;
                    sq = fmaf(d, d, sq);
                }
                const float rstd = 1.0f / sqrtf(warp_sum_f(sq) * (1.0f / 256.0f) + 1e-5f);
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.ds_lnw + 8 * c.lane)), w1 = __ldg(reinterpret_cast<const float4*>(p.ds_lnw + 8 * c.lane + 4));
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ds_lnb + 8 * c.lane)), g1 = __ldg(reinterpret_cast<const float4*>(p.ds_lnb + 8 * c.lane + 4));
                const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w}, bb[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) v8[i] = gelu_erf((v8[i] - mean) * rstd * ww[i] + bb[i]);
                const float4 o0 = make_float4(v8[0], v8[1], v8[2], v8[3]), o1 = make_float4(v8[4], v8[5], v8[6], v8[7]);
                float* dst[3] = {p.ring + (((size_t)id * 2 + ch) * p.T + (cnt - 1) % p.T) * kD, xo + (size_t)(c.t - 1) * kD,
                                 p.e_out ? p.e_out + (size_t)n * kD : nullptr};
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (dst[k]) {
                        *reinterpret_cast<float4*>(dst[k] + 8 * c.lane) = o0;
                        *reinterpret_cast<float4*>(dst[k] + 8 * c.lane + 4) = o1;
                    }
            }
            if (c.r == 0 && c.tid == 0) p.tvalid[c.b] = c.t;
        }
        cl_arrive();
        cl_wait();
        if (d2) d2[6] = clock64();   // cluster barrier passed
    }
    if (dbg) p.dbg[p.n_ops] = clock64();
}

__device__ __forceinline__ int group_size(int ns);

// =========================== TMA producer (warp 8), one op ahead of the others ===========================
__device__ __forceinline__ void tma_loop(Ctx& c, const FusedParams& p) {
    int wi = 0;                      // running W stage counter
    for (int oi = 0; oi < p.n_ops; ++oi) {
        c.oi = oi;
        // this warp runs ahead of the shared-memory copy of the op: it reads the fields it needs from global memory
        const OpShape sh = op_shape(c, p.ops[oi].f);
        long long* d2 = fine_stamps(p, oi);
        if (sh.kind == FOP_GEMM && elect_one()) {
            int w = wi;
            if (d2) d2[10] = clock64();            // TMA: first issue of this op
            // tile order = MMA order: subtiles go in GROUPS of 2 or 3 that are issued interleaved (see mma_gemm)
            const int G = group_size(sh.ns);
            for (int kc = 0; kc < sh.kch; ++kc)
                for (int sp = 0; sp < sh.ns; sp += G)
                    for (int kb = 0; kb < 4; ++kb)
                        for (int sub = 0; sub < G; ++sub, ++w) {
                            const int s = w % kWStages;
                            const uint32_t ph = (uint32_t)(w / kWStages) & 1u;
                            fwait(c.w_empty(s), ph ^ 1u, 3, c.oi);
                            mbar_arrive_expect_tx(c.w_full(s), 2u * kWTile);
                            tma_load_2d(c.w_hi(s), &p.ops[oi].map_hi, kc * 256 + kb * kBK, sh.n_begin + (sp + sub) * 64, c.w_full(s));
                            tma_load_2d(c.w_lo(s), &p.ops[oi].map_lo, kc * 256 + kb * kBK, sh.n_begin + (sp + sub) * 64, c.w_full(s));
                        }
            if (d2) d2[11] = clock64();            // TMA: last issue
        }
        wi += sh.ns * 4 * sh.kch;
        __syncwarp();
        if (oi > 0) cl_wait();
        cl_arrive();
    }
    cl_wait();                       // this warp is one barrier behind
}

// ---- GEMM, MMA side (one elected thread).  Dependent MMAs on ONE accumulator issue at most every ~90 cycles whatever
// their width, while a 128 x 64 x 16 MMA is 32 cycles of tensor work: G subtiles (G accumulators) are issued interleaved
// so that the pipe is not idle between dependent MMAs.  G = 3 was measured SLOWER than G = 2 for the 6-subtile ops
// (32.5k vs 29.5k cycles): with 4 accumulator slots the second group of 3 waits for the epilogue of the first.
__device__ __forceinline__ int group_size(int ns) { return 2; }

template <int G>
__device__ __forceinline__ void mma_gemm(const Ctx& c, const OpShape& sh, int gst, int ga, int wi, long long* d2) {
    const uint32_t tmb = c.tmem_base;
    constexpr uint32_t idesc = make_idesc(64);
    int w = wi;
    for (int kc = 0; kc < sh.kch; ++kc) {
        for (int sp = 0; sp < sh.ns; sp += G) {
            uint32_t tm_acc[G];
#pragma unroll
            for (int u = 0; u < G; ++u) {
                const int g = gst + sp + u, slot = g & (kAccSlots - 1);
                tm_acc[u] = tmb + kTmAcc + (uint32_t)(slot * 64);
                if (kc == 0) fwait(c.acc_empty(slot), ((uint32_t)(g / kAccSlots) & 1u) ^ 1u, 4, c.oi);
            }
            tc_fence_after();
            for (int kb = 0; kb < 4; ++kb, w += G) {
                if (sp == 0) fwait(c.a_full(kb), (uint32_t)(ga + kc) & 1u, 5, c.oi);
                if (d2 && kc == 0 && sp == 0 && kb == 0) d2[7] = clock64();      // MMA: A k-block 0 ready
                uint32_t wsm[G];
#pragma unroll
                for (int u = 0; u < G; ++u) {
                    const int s = (w + u) % kWStages;
                    wsm[u] = c.w_hi(s);
                    fwait(c.w_full(s), (uint32_t)((w + u) / kWStages) & 1u, 6, c.oi);
                }
                tc_fence_after();
                if (d2 && kc == 0 && sp == 0 && kb == 0) d2[8] = clock64();      // MMA: first W stages ready
#pragma unroll
                for (int k = 0; k < kBK / kUmmaK; ++k) {
                    const uint32_t koff = (uint32_t)k * kUmmaK * 2;
                    const uint32_t ah = tmb + (uint32_t)(kb * 32 + k * 8), al = ah + kTmALo;
                    const uint32_t first = (kc | kb | k) ? 1u : 0u;
#pragma unroll
                    for (int u = 0; u < G; ++u) umma_bf16_ta(tm_acc[u], al, make_desc(wsm[u] + koff), idesc, first);      // small terms first
#pragma unroll
                    for (int u = 0; u < G; ++u) umma_bf16_ta(tm_acc[u], ah, make_desc(wsm[u] + kWTile + koff), idesc, 1u);
#pragma unroll
                    for (int u = 0; u < G; ++u) umma_bf16_ta(tm_acc[u], ah, make_desc(wsm[u] + koff), idesc, 1u);
                }
#pragma unroll
                for (int u = 0; u < G; ++u) umma_commit(c.w_empty((w + u) % kWStages));
            }
            if (kc == sh.kch - 1) {
#pragma unroll
                for (int u = 0; u < G; ++u) umma_commit(c.acc_full((gst + sp + u) & (kAccSlots - 1)));
            }
        }
        umma_commit(c.a_empty());
    }
    if (d2) d2[9] = clock64();             // MMA: last issue
}

// =========================== MMA issuer (warp 9) ===========================
__device__ __forceinline__ void mma_loop(Ctx& c, const FusedParams& p) {
    int gst = 0, ga = 0, wi = 0, na = 0;
    for (int oi = 0; oi < p.n_ops; ++oi) {
        const FOpFields& op = c.opslot[oi & 1];
        c.oi = oi;
        if (oi + 1 < p.n_ops)        // fields of the next op -> the other shared-memory slot
            reinterpret_cast<uint32_t*>(&c.opslot[(oi + 1) & 1])[c.lane] = __ldg(reinterpret_cast<const uint32_t*>(&p.ops[oi + 1].f) + c.lane);
        const OpShape sh = op_shape(c, op);
        long long* d2 = fine_stamps(p, oi);
        if (sh.kind == FOP_GEMM && elect_one()) {
            if (group_size(sh.ns) == 3) mma_gemm<3>(c, sh, gst, ga, wi, d2);
            else mma_gemm<2>(c, sh, gst, ga, wi, d2);
        }
        if (sh.kind == FOP_ATTN && elect_one()) attention_mma(c, na);
        if (sh.kind == FOP_ATTN) na += att_rounds(c);
        gst += sh.ns;
        ga += sh.kch;
        wi += sh.ns * 4 * sh.kch;
        __syncwarp();
        cl_arrive();
        cl_wait();
    }
}

// Warp-specialised register budgets (setmaxnreg, per warpgroup): the kernel launches with 168 registers per thread
// (12 warps); the two worker warpgroups grow to 208 (they hold a 128 x 256 fp32 tile in registers), the TMA / MMA /
// idle warpgroup shrinks to 88 (the pool only holds what the shrinking warps release: 8 x 40 = 4 x 80).  Each role runs its own copy of the op loop so that no code is shared between budgets.
#ifndef VAPB_F_NO_SETMAXNREG
__device__ __forceinline__ void reg_grow() { asm volatile("setmaxnreg.inc.sync.aligned.u32 208;"); }
__device__ __forceinline__ void reg_shrink() { asm volatile("setmaxnreg.dec.sync.aligned.u32 88;"); }
#else
__device__ __forceinline__ void reg_grow() {}
__device__ __forceinline__ void reg_shrink() {}
#endif

__global__ void __launch_bounds__(kThreadsF, 1) k_stream_tf(const FusedParams p) {
    extern __shared__ uint8_t smem_raw[];
    Ctx c;
    c.wbase = (smem_u32(smem_raw) + 1023u) & ~1023u;       // W ring first: SWIZZLE_128B tiles need 1024 B alignment
    c.uni = reinterpret_cast<float*>(smem_raw + (c.wbase - smem_u32(smem_raw)) + kWStages * 2 * kWTile);
    c.bars = c.wbase + kWStages * 2 * kWTile + kUniBytes;
    c.opslot = reinterpret_cast<FOpFields*>(reinterpret_cast<uint8_t*>(c.uni) + kUniBytes + 256);
    c.tid = threadIdx.x;
    c.warp = c.tid >> 5;
    c.lane = c.tid & 31;
    c.r = (int)cluster_ctarank();
    c.b = blockIdx.x >> 1;
    c.T = p.T;
    c.mode = p.mode;
    c.oi = 0;
    if (p.mode == 0) { c.m0 = c.b * 2 * p.T; c.rows = 2 * p.T; }
    else { c.m0 = (2 * c.b + c.r) * p.T; c.rows = p.T; }
    const int id = __ldg(p.ids + c.b);
    const int cnt = __ldg(p.count + id) + 1;           // frames including the one appended this step
    c.t = cnt < p.T ? cnt : p.T;

    if (c.warp == kWorkers && c.lane == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(c.a_full(i), kWorkers);
        mbar_init(c.a_empty(), 1);
        for (int i = 0; i < kWStages; ++i) {
            mbar_init(c.w_full(i), 1);
            mbar_init(c.w_empty(i), 1);
        }
        for (int i = 0; i < kAccSlots; ++i) {
            mbar_init(c.acc_full(i), 1);
            mbar_init(c.acc_empty(i), kWorkers);
        }
        mbar_init(c.att_in(), kWorkers);
        mbar_init(c.s_full(), 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(c.p_ready(i), 4);
            mbar_init(c.o_full(i), 1);
        }
        fence_barrier_init();
    }
    if (c.warp == kWorkers + 1) tmem_alloc(c.tmem_slot(), 512u);
    if (c.warp == 0) reinterpret_cast<uint32_t*>(c.opslot)[c.lane] = __ldg(reinterpret_cast<const uint32_t*>(&p.ops[0].f) + c.lane);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c.tmem_base) : "r"(c.tmem_slot()));

    if (c.warp < kWorkers) {
        reg_grow();
        workers_loop(c, p, id, cnt);
    } else {
        reg_shrink();
        if (c.warp == kWorkers) tma_loop(c, p);
        else if (c.warp == kWorkers + 1) mma_loop(c, p);
        else {
            // two spare warps complete the third warpgroup; the first one runs the small side tasks next to the op that
            // allows it (they only read X, which that op does not modify), so they cost no cluster barrier of their own;
            // the second one fetches every op's tensor maps up front (the op program is cold when a step starts)
            if (c.warp == kWorkers + 3)
                for (int i = c.lane; i < p.n_ops * 2; i += 32) prefetch_tmap(i & 1 ? &p.ops[i >> 1].map_lo : &p.ops[i >> 1].map_hi);
            for (int oi = 0; oi < p.n_ops; ++oi) {
                const int side = c.warp == kWorkers + 2 ? __ldg(&p.ops[oi].f.side) : 0;
                if (side == FSIDE_VAD) {
                    // vad = sigmoid(va_classifier(x[t-1])) on the ar_channel output (vap_main.py:292-293, 313-314)
                    const float* xr = p.X + ((size_t)(2 * c.b + c.r) * p.T + (c.t - 1)) * kD + 8 * c.lane;
                    const float4 x0 = ldcg4(xr), x1 = ldcg4(xr + 4);
                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.va_w + 8 * c.lane));
                    const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.va_w + 8 * c.lane + 4));
                    float s = 0.f;
                    s = fmaf(x0.x, w0.x, s); s = fmaf(x0.y, w0.y, s); s = fmaf(x0.z, w0.z, s); s = fmaf(x0.w, w0.w, s);
                    s = fmaf(x1.x, w1.x, s); s = fmaf(x1.y, w1.y, s); s = fmaf(x1.z, w1.z, s); s = fmaf(x1.w, w1.w, s);
                    s = warp_sum_f(s) + __ldg(p.va_b);
                    if (c.lane == 0) (p.io ? p.io->out : p.out)[c.b * 6 + 4 + c.r] = 1.0f / (1.0f + expf(-s));
                } else if (side == FSIDE_GATHER_LAST) {
                    // newest frame of channel r -> one row per sequence for the pruned layer's tail
                    const int n = 2 * c.b + c.r;
                    const float* xr = p.X + ((size_t)n * p.T + (c.t - 1)) * kD + 8 * c.lane;
                    float* xo = p.Xl + (size_t)n * kD + 8 * c.lane;
                    *reinterpret_cast<float4*>(xo) = ldcg4(xr);
                    *reinterpret_cast<float4*>(xo + 4) = ldcg4(xr + 4);
                }
                __syncwarp();
                cl_arrive();
                cl_wait();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (c.warp == kWorkers + 1) tmem_dealloc(c.tmem_base, 512u);
}

}  // namespace

size_t fused_smem_bytes() { return kSmemF; }

bool fused_prepare(std::string& err) {
    static OncePerDevice once;
    if (!once.first()) return true;
    cudaError_t e = cudaFuncSetAttribute(k_stream_tf, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemF);
    if (e != cudaSuccess) {
        err = std::string("cudaFuncSetAttribute(k_stream_tf) failed: ") + cudaGetErrorString(e);
        return false;
    }
    return true;
}

cudaError_t launch_fused_tf(const FusedParams& p, int B, cudaStream_t st) {
    static const int use_pdl = getenv("VAPB_STREAM_PDL") ? atoi(getenv("VAPB_STREAM_PDL")) : 0;     // measured 3.5 us SLOWER than a plain graph edge (636.2 vs 632.6 us per step): off
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * B);
    cfg.blockDim = dim3(kThreadsF);
    cfg.dynamicSmemBytes = kSmemF;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = use_pdl ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, k_stream_tf, p);
}

}  // namespace vapb
