// Per-stream persistent transformer kernel, second generation -- see fused_tf2.cuh for the scheme.
//
// Cluster of FOUR CTAs = two streams x two column halves (rank k: stream slot k >> 1, half r = k & 1).  Every CTA needs its
// half of every weight matrix for every stream, so both operand streams are multicast (this halves the L2 -> SM traffic;
// it did not change the run time: the fabric is not the limiter): the two CTAs of a stream share every A tile (each issues 64 of its 128 rows to both), the two CTAs with
// the same half r share every W tile across the two streams (each issues 64 of its 128 rows to both).  "Empty" barriers
// therefore collect one tcgen05.commit from each of the two consumers.  An odd batch gets a ghost stream that runs the
// same pipeline on scratch rows and writes no state.
//
// CTA anatomy (384 threads, 1 CTA per SM):
//   warps 0-7  workers: GEMM epilogues (tcgen05.ld -> LayerNorm correction / GELU / residual -> bf16 hi / lo planes ->
//              staging tile -> TMA store), softmax, ring gather
//   warp 8     TMA producer of the W tiles; weights are constants, so it runs AHEAD of the op barriers (throttled only by
//              the ring): the first W tiles of an op are in shared memory before the barrier in front of it opens
//   warp 9     TMEM owner + the single thread that issues tcgen05.mma
//   warp 10    TMA producer of everything that depends on the previous op: A tiles; Q / K / V tiles of the attention
//              (warps 9 and 10 prefetch every op's tensor maps while the ring is gathered)
//   warp 11    side tasks (newest embedding of the ring gather, vad, newest-frame gather)
// Every stage holds the hi and the lo plane of one 128 x 64 bf16 tile (2 x 16 KB, 128 B swizzle): an A ring (A tiles
// stream once per subtile group; Q of the attention) and a W ring (W tiles; K and V of the attention at reserved ring
// positions).  Each worker warp owns a 4 KB staging tile through which every global access of the epilogue passes:
// tensor memory hands a thread one ROW, but a warp-wide access that touches 32 different 128 B lines costs the LSU 32
// passes (measured: thread-per-row stores made the first cut of this kernel slower than its predecessor), so the rows
// are written into the tile in the TMA box layout and leave as TMA stores.
// All pipelines carry their phase across ops through running counters that every role advances identically.
#include "fused_tf2.cuh"
#include "tc_ptx.cuh"

#include <cstdlib>
#include <string>

namespace vapb {

namespace {

using namespace tcp;

constexpr int kWorkers2 = 8;
constexpr int kThreads2 = (kWorkers2 + 4) * 32;
constexpr int kPlane = 128 * kBK * 2;                 // one bf16 plane of a 128 x 64 tile: 16 KB
constexpr int kStage = 2 * kPlane;                    // hi + lo
#ifndef VAPB_F2_ASTAGES
#define VAPB_F2_ASTAGES 2
#endif
#ifndef VAPB_F2_WSTAGES
#define VAPB_F2_WSTAGES 4
#endif
constexpr int kAStages = VAPB_F2_ASTAGES, kWStg = VAPB_F2_WSTAGES, kAcc = 4;      // 4 accumulators of 128 columns = all of tensor memory
constexpr int kStgBytes = 4096;                       // per worker warp: 32 rows x 128 bytes
constexpr int kSmem2 = (kAStages + kWStg) * kStage + kWorkers2 * kStgBytes + 1024 /*alignment slack*/ + 512 /*barriers, op slots*/;
static_assert(kAStages >= 2 && (kAStages % 2) == 0 && kWStg >= 3, "ring sizes");
static_assert(kSmem2 <= 232448, "shared memory budget");

__device__ __forceinline__ void cl_arrive2() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cl_wait2() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
// generic-proxy global writes (epilogue stores) -> async-proxy reads (TMA) of this or the peer CTA
__device__ __forceinline__ void fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ float warp_sum2(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __noinline__ __attribute__((noreturn)) void f2wait_fail(int tag, int oi, uint32_t parity) {
    if ((threadIdx.x & 31) == 0)
        printf("vapb stream kernel v2: wait %d (1 a_full 2 a_empty 3 w_full 4 w_empty 5 acc_full 6 acc_empty 7 s_full 8 p_ready 9 o_full) timed out, op %d block %d warp %d parity %u\n",
               tag, oi, blockIdx.x, threadIdx.x >> 5, parity);
    __trap();
    for (;;) {}
}
__device__ __forceinline__ void f2wait(uint32_t bar, uint32_t parity, int tag, int oi) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (VAPB_WAIT_POLLS != 0u && ++polls > VAPB_WAIT_POLLS) f2wait_fail(tag, oi, parity);
    }
}

struct Ctx2 {
    long long* fine;       // fine clock stamps of one op (cluster 0 / CTA 0 only), else null
    uint32_t sbase, bars, tmem_base;
    F2Fields* opslot;
    int tid, warp, lane;
    int b, r;              // stream index in the batch (= workspace slot), column half
    int sidx, ghost;       // stream slot inside the cluster; ghost = padding stream of an odd batch (no state is written)
    uint16_t mask_a, mask_w;   // multicast masks: the two CTAs of this stream / the two CTAs of this column half
    int T, t, oi;
    __device__ __forceinline__ uint32_t a_stage(int s) const { return sbase + (uint32_t)s * kStage; }
    __device__ __forceinline__ uint32_t w_stage(int s) const { return sbase + (uint32_t)(kAStages + s) * kStage; }
    __device__ __forceinline__ uint32_t stg() const { return sbase + (uint32_t)(kAStages + kWStg) * kStage + (uint32_t)warp * kStgBytes; }
    __device__ __forceinline__ uint32_t a_full(int s) const { return bars + 8u * s; }
    __device__ __forceinline__ uint32_t a_empty(int s) const { return bars + 8u * (kAStages + s); }
    __device__ __forceinline__ uint32_t w_full(int s) const { return bars + 16u * kAStages + 8u * s; }
    __device__ __forceinline__ uint32_t w_empty(int s) const { return bars + 16u * kAStages + 8u * (kWStg + s); }
    __device__ __forceinline__ uint32_t misc() const { return bars + 16u * (kAStages + kWStg); }
    __device__ __forceinline__ uint32_t acc_full(int s) const { return misc() + 8u * s; }
    __device__ __forceinline__ uint32_t acc_empty(int s) const { return misc() + 32u + 8u * s; }
    __device__ __forceinline__ uint32_t s_full() const { return misc() + 64u; }
    __device__ __forceinline__ uint32_t p_ready(int x) const { return misc() + 72u + 8u * x; }
    __device__ __forceinline__ uint32_t o_full(int x) const { return misc() + 88u + 8u * x; }
    __device__ __forceinline__ uint32_t tmem_slot() const { return misc() + 104u; }
};


// timing experiments only (results are wrong): VAPB_F2_NOSTORE drops every epilogue store, _NOSTORE_P only the plane
// stores, _NOSTORE_X only the fp32 stores; VAPB_F2_NOFENCE drops the proxy fence in front of the op barrier; VAPB_F2_NOGELU
#if defined(VAPB_F2_NOSTORE) || defined(VAPB_F2_NOSTORE_X)
#define F2_STORE(x)
#else
#define F2_STORE(x) x
#endif
#if defined(VAPB_F2_NOSTORE) || defined(VAPB_F2_NOSTORE_P)
#define F2_STORE_P(x)
#else
#define F2_STORE_P(x) x
#endif

// ---- per-warp staging tile: 32 rows x 128 bytes, 16-byte chunks XOR-swizzled by the row (conflict-free both ways) ----
__device__ __forceinline__ uint32_t stg_addr(uint32_t stg, int row, int chunk) { return stg + (uint32_t)row * 128u + ((uint32_t)(chunk ^ (row & 7)) << 4); }
// thread-per-row side: lane = row, 32 words
__device__ __forceinline__ void stg_put_row(uint32_t stg, int lane, const uint32_t* w) {
#pragma unroll
    for (int ch = 0; ch < 8; ++ch)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_addr(stg, lane, ch)), "r"(w[4 * ch]), "r"(w[4 * ch + 1]), "r"(w[4 * ch + 2]),
                     "r"(w[4 * ch + 3])
                     : "memory");
}
__device__ __forceinline__ void stg_get_row(uint32_t stg, int lane, uint32_t* w) {
#pragma unroll
    for (int ch = 0; ch < 8; ++ch)
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4 * ch]), "=r"(w[4 * ch + 1]), "=r"(w[4 * ch + 2]), "=r"(w[4 * ch + 3])
                     : "r"(stg_addr(stg, lane, ch))
                     : "memory");
}
// coalesced side: pass i covers rows 4i .. 4i + 3, 8 lanes x 16 bytes = the 128 bytes of one row
__device__ __forceinline__ void stg_put_co(uint32_t stg, int lane, int i, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_addr(stg, 4 * i + (lane >> 3), lane & 7)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
// staging tile (32 rows x 128 bytes, exactly the SWIZZLE_128B box layout) -> global through ONE TMA store issued by lane 0.
// The LSU never sees these bytes: thread-per-row st.global made the first cut of this kernel slower than its predecessor,
// coalesced st.global through the tile still cost ~150 k cycles per launch (timing-only builds without the stores).
// WAIT: returns when the tile may be rewritten (the TMA engine has read it); otherwise the caller runs stg_wait_read()
// before it touches the tile again, with independent work in between (-16 k cycles per launch against waiting in place).
// Completion of the writes is awaited once per op.  Measured and not kept: sending the last subtile's second plane through
// tiles borrowed from the idle A ring so that no read-wait is left at the end of an op (+12 k cycles: the kernel is
// sensitive to code size), see profiles/r02_r_*.
#ifdef VAPB_F2_EAGERWAIT
#define F2_EAGER true
#else
#define F2_EAGER false
#endif
// Stores address [sequence][position][column]: the 32 rows of warp quadrant q of stream slot b are positions 32 (q & 1) ..
// of sequence 2 b + (q >> 1); the maps end at position T, so the padding rows of a tile are clipped.
template <bool PLANE, bool WAIT = true>
__device__ __forceinline__ void stg_tma_store_2d(uint32_t stg, int lane, const CUtensorMap* map, int c0, int pos0, int seq) {
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        if (PLANE) { F2_STORE_P(tma_store_3d(map, c0, pos0, seq, stg);) } else { F2_STORE(tma_store_3d(map, c0, pos0, seq, stg);) }
        tma_store_commit();
        if (WAIT || F2_EAGER) tma_store_wait_read();
    }
    __syncwarp();
}
__device__ __forceinline__ void stg_wait_read(int lane) {
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
}
// end of an op: every store this warp issued is complete before the op barrier publishes the results
__device__ __forceinline__ void stores_done(int lane) {
    if (lane == 0) tma_store_wait_all();
    __syncwarp();
}

// mean and M2 (sum of squared deviations) of 32 values, two passes in registers
__device__ __forceinline__ float2 block_stats32(const float (&v)[32]) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int e = 0; e < 32; e += 2) { s0 += v[e]; s1 += v[e + 1]; }
    const float mean = (s0 + s1) * (1.0f / 32.0f);
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int e = 0; e < 32; e += 2) {
        const float d0 = v[e] - mean, d1 = v[e + 1] - mean;
        q0 = fmaf(d0, d0, q0);
        q1 = fmaf(d1, d1, q1);
    }
    return make_float2(mean, q0 + q1);
}
// LayerNorm statistics of a 256-wide row from its 8 block partials (exact pairwise combination)
__device__ __forceinline__ void row_stats(const float* st_row, float& mu, float& rstd) {
    float4 p[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = __ldcg(reinterpret_cast<const float4*>(st_row) + i);
    const float m[8] = {p[0].x, p[0].z, p[1].x, p[1].z, p[2].x, p[2].z, p[3].x, p[3].z};
    const float q[8] = {p[0].y, p[0].w, p[1].y, p[1].w, p[2].y, p[2].w, p[3].y, p[3].w};
    float sm = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sm += m[i];
    mu = sm * 0.125f;
    float m2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float d = m[i] - mu;
        m2 += q[i] + 32.0f * d * d;
    }
    rstd = 1.0f / sqrtf(m2 * (1.0f / 256.0f) + 1e-5f);
}

// ============================== GEMM: worker side (epilogue only) ==============================
// Warp (q = warp & 3, hf = warp >> 2) drains columns [64 hf, 64 hf + 64) of every 128-column subtile for the 32 rows
// of TMEM lane quadrant q.  Everything that goes to or comes from global memory passes through the warp's staging tile.
__device__ __forceinline__ void ln_correct(float (&v)[32], const F2Fields& op, int col, float mu, float rstd) {
    const float nm = -mu * rstd;
#pragma unroll
    for (int e4 = 0; e4 < 8; ++e4) {
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(op.ln_s + col) + e4);
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(op.ln_c + col) + e4);
        v[4 * e4] = fmaf(rstd, v[4 * e4], fmaf(nm, s4.x, c4.x));
        v[4 * e4 + 1] = fmaf(rstd, v[4 * e4 + 1], fmaf(nm, s4.y, c4.y));
        v[4 * e4 + 2] = fmaf(rstd, v[4 * e4 + 2], fmaf(nm, s4.z, c4.z));
        v[4 * e4 + 3] = fmaf(rstd, v[4 * e4 + 3], fmaf(nm, s4.w, c4.w));
    }
}

// accumulator block (32 columns of this thread's row) -> registers; the warp's second block releases the accumulator
__device__ __forceinline__ void load_acc_block(const Ctx2& c, uint32_t taddr, int slot, bool last, float (&v)[32]) {
    uint32_t raw[32];
    tmem_ld32(taddr, raw);
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(raw[e]);
    if (last) {
        tc_fence_before();
        __syncwarp();
        if (c.lane == 0) mbar_arrive(c.acc_empty(slot));
    }
}

// planes out (+ LayerNorm correction, + GELU).  One 32-column block at a time, ROLLED (the kernel is instruction-cache
// sensitive): the block's hi / lo halves (64 bytes per row and plane) go through the two halves of the warp's staging tile
// (SWIZZLE_64B boxes) and leave as two TMA stores; the engine reads them while the next block is loaded, corrected,
// activated and split.
__device__ __forceinline__ uint32_t stgh_addr(uint32_t base, int row, int chunk) { return base + (uint32_t)row * 64u + ((uint32_t)(chunk ^ ((row >> 1) & 3)) << 4); }
__device__ __forceinline__ void stgh_put_row(uint32_t base, int lane, const uint32_t* w) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stgh_addr(base, lane, ch)), "r"(w[4 * ch]), "r"(w[4 * ch + 1]), "r"(w[4 * ch + 2]),
                     "r"(w[4 * ch + 3])
                     : "memory");
}
// hi / lo halves of a 32 x 32 block -> the two halves of the warp's staging tile -> two TMA stores (no wait: the caller runs
// stg_wait_read() before it touches the tile again)
__device__ __forceinline__ void store_block_halves(uint32_t stg, int lane, const uint32_t* H, const uint32_t* L, const CUtensorMap* map_h,
                                                   const CUtensorMap* map_l, int col, int pos0, int seq) {
    stgh_put_row(stg, lane, H);
    stgh_put_row(stg + kStgBytes / 2, lane, L);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        F2_STORE_P(tma_store_3d(map_h, col, pos0, seq, stg);)
        F2_STORE_P(tma_store_3d(map_l, col, pos0, seq, stg + kStgBytes / 2);)
        tma_store_commit();
    }
    __syncwarp();
}
__device__ __noinline__ void epilogue_planes(const Ctx2& c, const F2Op* gop, const F2Fields& op, const Fused2Params& p, int gs) {
    const int q = c.warp & 3, hf = c.warp >> 2;
    const size_t grow = (size_t)c.b * 128 + 32 * q + c.lane;
    const int ns = op.N >> 8;
    const uint32_t tm_row = c.tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t stg = c.stg();
    float mu = 0.f, rstd = 1.f;
    if (op.n_ln > 0) row_stats(p.stats + grow * 16, mu, rstd);
    long long* fine = c.tid == 0 ? c.fine : nullptr;
    for (int s = 0; s < ns; ++s) {
        const int g = gs + s, slot = g & (kAcc - 1);
        const int col0 = 256 * s + 128 * c.r + 64 * hf;
        f2wait(c.acc_full(slot), (uint32_t)(g / kAcc) & 1u, 5, c.oi);
        tc_fence_after();
        if (fine && s == 0) fine[1] = clock64();
        if (fine && s == ns - 1) fine[2] = clock64();
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
            const int col = col0 + 32 * blk;
            float v[32];
            load_acc_block(c, tm_row + (uint32_t)(slot * 128 + 64 * hf + 32 * blk), slot, blk == 1, v);
            if (col < op.n_ln) ln_correct(v, op, col, mu, rstd);
#ifndef VAPB_F2_NOGELU
            if (op.act == 1) gelu_block(v);          // eight independent chains in flight (tc_ptx.cuh)
#endif
            uint32_t H[16], L[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) split2(v[2 * e], v[2 * e + 1], H[e], L[e]);
            stg_wait_read(c.lane);                   // the previous block's halves: read while this one was computed
            store_block_halves(stg, c.lane, H, L, &gop->m[4], &gop->m[5], col, 32 * (q & 1), 2 * c.b + (q >> 1));
        }
    }
    stores_done(c.lane);
    if (fine) fine[3] = clock64();
}

// X update: x += acc (residual), fp32 rows + planes + LayerNorm block statistics.  The residual block is fetched coalesced
// one block ahead and passes through the staging tile to reach the thread that owns the row; results leave through TMA.
__device__ __noinline__ void epilogue_x(const Ctx2& c, const F2Op* gop, const F2Fields& op, const Fused2Params& p, int gs) {
    const int q = c.warp & 3, hf = c.warp >> 2;
    const size_t grow = (size_t)c.b * 128 + 32 * q + c.lane;
    const int grow0 = c.b * 128 + 32 * q;
    const int ns = op.N >> 8;
    const uint32_t tm_row = c.tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t stg = c.stg();
    long long* fine = c.tid == 0 ? c.fine : nullptr;
    auto res_ptr = [&](int col, int i) { return reinterpret_cast<const uint4*>(p.Xf + ((size_t)grow0 + 4 * i + (c.lane >> 3)) * kD + col + 4 * (c.lane & 7)); };
    for (int s = 0; s < ns; ++s) {
        const int g = gs + s, slot = g & (kAcc - 1);
        const int col0 = 256 * s + 128 * c.r + 64 * hf;
        uint4 rc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rc[i] = __ldcg(res_ptr(col0, i));
        f2wait(c.acc_full(slot), (uint32_t)(g / kAcc) & 1u, 5, c.oi);
        tc_fence_after();
        if (fine && s == 0) fine[1] = clock64();
        if (fine && s == ns - 1) fine[2] = clock64();
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
            const int col = col0 + 32 * blk;
            stg_wait_read(c.lane);                              // the stores issued before this point have left the tile
#pragma unroll
            for (int i = 0; i < 8; ++i) stg_put_co(stg, c.lane, i, rc[i]);
            __syncwarp();
            float v[32];
            {
                uint32_t xr[32];
                stg_get_row(stg, c.lane, xr);
                __syncwarp();
                if (blk == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) rc[i] = __ldcg(res_ptr(col0 + 32, i));       // next block's residual in flight
                }
                load_acc_block(c, tm_row + (uint32_t)(slot * 128 + 64 * hf + 32 * blk), slot, blk == 1, v);
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] += __uint_as_float(xr[e]);
            }
            const float2 st = block_stats32(v);
            F2_STORE(*reinterpret_cast<float2*>(p.stats + grow * 16 + (col >> 5) * 2) = st;)
            {
                uint32_t vb[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) vb[e] = __float_as_uint(v[e]);
                stg_put_row(stg, c.lane, vb);
            }
            stg_tma_store_2d<false, false>(stg, c.lane, &gop->m[6], col, 32 * (q & 1), 2 * c.b + (q >> 1));   // fp32 residual stream (read while the planes are split)
            uint32_t H[16], L[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) split2(v[2 * e], v[2 * e + 1], H[e], L[e]);
            stg_wait_read(c.lane);
            store_block_halves(stg, c.lane, H, L, &gop->m[4], &gop->m[5], col, 32 * (q & 1), 2 * c.b + (q >> 1));
        }
    }
    stores_done(c.lane);
    if (fine) fine[3] = clock64();
}

// fp32 rows in the batched kernels' layout (K / V of the pruned layer for the newest-frame tail)
__device__ __noinline__ void epilogue_f32(const Ctx2& c, const F2Op* gop, const F2Fields& op, const Fused2Params& p, int gs) {
    const int q = c.warp & 3, hf = c.warp >> 2;
    const size_t grow = (size_t)c.b * 128 + 32 * q + c.lane;
    const int ns = op.N >> 8;
    const uint32_t tm_row = c.tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t stg = c.stg();
    float mu = 0.f, rstd = 1.f;
    if (op.n_ln > 0) row_stats(p.stats + grow * 16, mu, rstd);
    long long* fine = c.tid == 0 ? c.fine : nullptr;
    for (int s = 0; s < ns; ++s) {
        const int g = gs + s, slot = g & (kAcc - 1);
        const int col0 = 256 * s + 128 * c.r + 64 * hf;
        f2wait(c.acc_full(slot), (uint32_t)(g / kAcc) & 1u, 5, c.oi);
        tc_fence_after();
        if (fine && s == 0) fine[1] = clock64();
        if (fine && s == ns - 1) fine[2] = clock64();
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
            const int col = col0 + 32 * blk;
            float v[32];
            load_acc_block(c, tm_row + (uint32_t)(slot * 128 + 64 * hf + 32 * blk), slot, blk == 1, v);
            if (col < op.n_ln) ln_correct(v, op, col, mu, rstd);
            uint32_t vb[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) vb[e] = __float_as_uint(v[e]);
            stg_wait_read(c.lane);                              // the previous block's store: read during this block's load + correction
            stg_put_row(stg, c.lane, vb);
            // rows 32q .. 32q + 31 of the tile = positions 32 (q & 1) .. + 31 of sequence q >> 1; positions >= T are clipped by the map
            if (!c.ghost) {
                fence_proxy_async();
                __syncwarp();
                if (c.lane == 0) {
                    F2_STORE(tma_store_3d(&gop->m[col < 512 ? 4 : 5], col & 511, 32 * (q & 1), 2 * c.b + (q >> 1), stg);)
                    tma_store_commit();
                    if (F2_EAGER) tma_store_wait_read();
                }
            }
            __syncwarp();
        }
    }
    stores_done(c.lane);
    if (fine) fine[3] = clock64();
}

__device__ __forceinline__ void gemm_epilogue(const Ctx2& c, const F2Fields& op, const Fused2Params& p, int gs) {
    const int mode = __shfl_sync(0xffffffffu, op.out_mode, 0);
    const F2Op* gop = &p.ops[c.oi];
    if (mode == F2_OUT_PLANES) epilogue_planes(c, gop, op, p, gs);
    else if (mode == F2_OUT_X) epilogue_x(c, gop, op, p, gs);
    else epilogue_f32(c, gop, op, p, gs);
}

// ============================== GEMM: tile schedule ==============================
// Subtiles go in groups: a PAIR of subtiles (two W tiles per k-block, two 128-column accumulators issued interleaved)
// or a SINGLE subtile (one W tile per k-block, one chain of N = 128 MMAs).  The A tiles stream once per group.
__device__ __forceinline__ int gemm_groups(const F2Fields& op) { return ((op.N >> 8) + 1) >> 1; }
__device__ __forceinline__ int gemm_a_tiles(const F2Fields& op) { return (op.K >> 6) * gemm_groups(op); }     // even (K/64 is 4 or 12)
__device__ __forceinline__ int gemm_w_tiles(const F2Fields& op) { return (op.N >> 8) * (op.K >> 6); }

// W tiles (warp 8, ahead of the op barriers)
__device__ __forceinline__ void gemm_tma_w(const Ctx2& c, const F2Op* gop, int N, int K, int wt) {
    const int ns = N >> 8, nkb = K >> 6;
    int w = wt;
    if (c.fine) c.fine[7] = clock64();                              // W TMA: first issue
    for (int s = 0; s < ns; s += 2) {
        const int gsz = (s + 1 < ns) ? 2 : 1;
        for (int kb = 0; kb < nkb; ++kb)
            for (int u = 0; u < gsz; ++u, ++w) {
                const int st = w % kWStg;
                f2wait(c.w_empty(st), ((uint32_t)(w / kWStg) & 1u) ^ 1u, 4, c.oi);
                mbar_arrive_expect_tx(c.w_full(st), kStage);
                // this CTA's share of the tile: 64 of its 128 rows, delivered to both CTAs that compute column half r
                const int n0 = 256 * (s + u) + 128 * c.r + 64 * c.sidx;
                const uint32_t off = (uint32_t)c.sidx * (kPlane / 2);
                tma_load_2d_mc(c.w_stage(st) + off, &gop->m[2], kb * kBK, n0, c.w_full(st), c.mask_w);
                tma_load_2d_mc(c.w_stage(st) + kPlane + off, &gop->m[3], kb * kBK, n0, c.w_full(st), c.mask_w);
#ifdef VAPB_F2_TRACE
                if (blockIdx.x < 2 && c.oi < 3) printf("  W tile issued block %d op %d w %d stage %d n0 %d kb %d\n", blockIdx.x, c.oi, w, st, n0, kb);
#endif
            }
    }
    if (c.fine) c.fine[8] = clock64();                              // W TMA: last issue
}
// A tiles (warp 10, after the op barrier)
__device__ __forceinline__ void gemm_tma_a(const Ctx2& c, const F2Op* gop, const F2Fields& op, int at) {
    const int nkb = op.K >> 6, ng = gemm_groups(op);
    const int row0 = c.b * 128;
    int ai = at;
    for (int gi = 0; gi < ng; ++gi)
        for (int kb = 0; kb < nkb; ++kb, ++ai) {
            const int st = ai % kAStages;
            f2wait(c.a_empty(st), ((uint32_t)(ai / kAStages) & 1u) ^ 1u, 2, c.oi);
            mbar_arrive_expect_tx(c.a_full(st), kStage);
            // this CTA's share: rows 64r .. 64r + 63 of the stream's tile, delivered to both CTAs of the stream
            const uint32_t off = (uint32_t)c.r * (kPlane / 2);
            tma_load_2d_mc(c.a_stage(st) + off, &gop->m[0], kb * kBK, row0 + 64 * c.r, c.a_full(st), c.mask_a);
            tma_load_2d_mc(c.a_stage(st) + kPlane + off, &gop->m[1], kb * kBK, row0 + 64 * c.r, c.a_full(st), c.mask_a);
        }
}

// ============================== GEMM: MMA side (one elected thread) ==============================
template <bool PAIR>
__device__ __forceinline__ void mma_group(const Ctx2& c, int nkb, bool first_group, int& ai, int& w, int slot0, int slot1) {
#ifdef VAPB_F2_SINGLE64
    constexpr bool kTwo = true;                           // experiment: single groups as two interleaved N = 64 halves
    constexpr uint32_t idesc = PAIR ? make_idesc(128) : make_idesc(64);
#else
    constexpr bool kTwo = PAIR;                           // a single group is ONE chain of N = 128 MMAs: the two-halves form read every A tile twice
    constexpr uint32_t idesc = make_idesc(128);
#endif
    const uint32_t acc0 = c.tmem_base + (uint32_t)(slot0 * 128);
    const uint32_t acc1 = PAIR ? c.tmem_base + (uint32_t)(slot1 * 128) : acc0 + 64u;
    for (int kb = 0; kb < nkb; ++kb, ++ai) {
        const int ast = ai % kAStages;
        f2wait(c.a_full(ast), (uint32_t)(ai / kAStages) & 1u, 1, c.oi);
        const int ws0 = w % kWStg;
        f2wait(c.w_full(ws0), (uint32_t)(w / kWStg) & 1u, 3, c.oi);
        uint32_t wb0 = c.w_stage(ws0), wb1;
        int ws1 = ws0;
        if (PAIR) {
            ws1 = (w + 1) % kWStg;
            f2wait(c.w_full(ws1), (uint32_t)((w + 1) / kWStg) & 1u, 3, c.oi);
            wb1 = c.w_stage(ws1);
        } else {
            wb1 = wb0 + 64u * 128u;                      // rows 64..127 of the same W tile
        }
        tc_fence_after();
        if (c.fine && first_group && kb == 0) c.fine[5] = clock64();     // MMA: first operands ready
        const uint32_t ab = c.a_stage(ast);
#pragma unroll
        for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint32_t koff = (uint32_t)k * kUmmaK * 2;
            const uint32_t first = (kb | k) ? 1u : 0u;
            // small terms first: lo * hi, hi * lo, then hi * hi; the two accumulators alternate
            umma_bf16(acc0, make_desc(ab + kPlane + koff), make_desc(wb0 + koff), idesc, first);
            if (kTwo) umma_bf16(acc1, make_desc(ab + kPlane + koff), make_desc(wb1 + koff), idesc, first);
            umma_bf16(acc0, make_desc(ab + koff), make_desc(wb0 + kPlane + koff), idesc, 1u);
            if (kTwo) umma_bf16(acc1, make_desc(ab + koff), make_desc(wb1 + kPlane + koff), idesc, 1u);
            umma_bf16(acc0, make_desc(ab + koff), make_desc(wb0 + koff), idesc, 1u);
            if (kTwo) umma_bf16(acc1, make_desc(ab + koff), make_desc(wb1 + koff), idesc, 1u);
        }
        umma_commit_mc(c.w_empty(ws0), c.mask_w);
        if (PAIR) umma_commit_mc(c.w_empty(ws1), c.mask_w);
        umma_commit_mc(c.a_empty(ast), c.mask_a);
        w += PAIR ? 2 : 1;
    }
    umma_commit(c.acc_full(slot0));
    if (PAIR) umma_commit(c.acc_full(slot1));
}

__device__ __forceinline__ void gemm_mma(const Ctx2& c, const F2Fields& op, int at, int wt, int gs) {
    const int ns = op.N >> 8, nkb = op.K >> 6;
    int w = wt, ai = at;
    for (int s = 0; s < ns; s += 2) {
        const bool pair = s + 1 < ns;
        const int g0 = gs + s, g1 = gs + s + 1;
        const int slot0 = g0 & (kAcc - 1), slot1 = g1 & (kAcc - 1);
        f2wait(c.acc_empty(slot0), ((uint32_t)(g0 / kAcc) & 1u) ^ 1u, 6, c.oi);
        if (pair) f2wait(c.acc_empty(slot1), ((uint32_t)(g1 / kAcc) & 1u) ^ 1u, 6, c.oi);
        tc_fence_after();
        if (pair) mma_group<true>(c, nkb, s == 0, ai, w, slot0, slot1);
        else mma_group<false>(c, nkb, s == 0, ai, w, slot0, slot1);
    }
    if (c.fine) c.fine[6] = clock64();                              // MMA: last issue
}

// ============================== attention (modules.py:82-110, 170-212) ==============================
// Two heads (2r, 2r + 1) per CTA.  Shared memory: Q of head x in the next two A ring stages; K of head x and V of head x
// in the next four W ring positions (K0 K1 V0 V1; the W producer skips them, warp 10 fills them after the op barrier).
// K / V tiles hold the keys of sequence c (or of its sibling for the cross attention) in rows 64c .. 64c + 63, so row
// (c, i) of Q always finds its keys in columns 64c + j of S.  Tensor memory: S_x / P_x at columns 128x, O_x at 256 + 64x.
constexpr uint32_t kTmS2 = 0, kTmO2 = 256;
__host__ __device__ constexpr uint32_t make_idesc_bmn2(int bn) { return make_idesc(bn) | (1u << 16); }   // B operand MN-major

__device__ __forceinline__ void attn_tma(const Ctx2& c, const F2Op* gop, const F2Fields& op, int at, int wt) {
    const int row0 = c.b * 128;
    for (int x = 0; x < 2; ++x) {                          // Q
        const int ai = at + x, st = ai % kAStages;
        f2wait(c.a_empty(st), ((uint32_t)(ai / kAStages) & 1u) ^ 1u, 2, c.oi);
        mbar_arrive_expect_tx(c.a_full(st), kStage);
        const int col = op.qcol + (2 * c.r + x) * 64;
        tma_load_2d(c.a_stage(st), &gop->m[0], col, row0, c.a_full(st));
        tma_load_2d(c.a_stage(st) + kPlane, &gop->m[1], col, row0, c.a_full(st));
    }
    for (int i = 0; i < 4; ++i) {                          // K0 K1 V0 V1: two boxes of 64 rows each (swapped for the sibling channel)
        const int w = wt + i, st = w % kWStg;
        f2wait(c.w_empty(st), ((uint32_t)(w / kWStg) & 1u) ^ 1u, 4, c.oi);
        mbar_arrive_expect_tx(c.w_full(st), kStage);
        const int col = (i < 2 ? op.kcol : op.vcol) + (2 * c.r + (i & 1)) * 64;
        for (int half = 0; half < 2; ++half) {
            const int src = row0 + 64 * (half ^ op.sibling);
            tma_load_2d(c.w_stage(st) + half * (kPlane / 2), &gop->m[2], col, src, c.w_full(st));
            tma_load_2d(c.w_stage(st) + kPlane + half * (kPlane / 2), &gop->m[3], col, src, c.w_full(st));
        }
    }
}

__device__ __forceinline__ void attn_mma(const Ctx2& c, int at, int wt, int na) {
    constexpr uint32_t idesc_s = make_idesc(128);
    constexpr uint32_t idesc_o = make_idesc_bmn2(64);
    const uint32_t par = (uint32_t)na & 1u;
    for (int x = 0; x < 2; ++x) {
        const int ai = at + x, w = wt + x;
        f2wait(c.a_full(ai % kAStages), (uint32_t)(ai / kAStages) & 1u, 1, c.oi);
        f2wait(c.w_full(w % kWStg), (uint32_t)(w / kWStg) & 1u, 3, c.oi);
    }
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t koff = (uint32_t)k * kUmmaK * 2;
#pragma unroll
        for (int prod = 0; prod < 3; ++prod) {
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                const uint32_t qb = c.a_stage((at + x) % kAStages) + koff, kb = c.w_stage((wt + x) % kWStg) + koff;
                const uint32_t acc = c.tmem_base + kTmS2 + 128u * x;
                if (prod == 0) umma_bf16(acc, make_desc(qb + kPlane), make_desc(kb), idesc_s, k ? 1u : 0u);
                else if (prod == 1) umma_bf16(acc, make_desc(qb), make_desc(kb + kPlane), idesc_s, 1u);
                else umma_bf16(acc, make_desc(qb), make_desc(kb), idesc_s, 1u);
            }
        }
    }
    umma_commit(c.s_full());
    for (int x = 0; x < 2; ++x) {                           // Q and K are free once S is complete
        umma_commit_mc(c.a_empty((at + x) % kAStages), c.mask_a);
        umma_commit_mc(c.w_empty((wt + x) % kWStg), c.mask_w);
    }
    f2wait(c.p_ready(0), par, 8, c.oi);
    f2wait(c.p_ready(1), par, 8, c.oi);
    for (int x = 0; x < 2; ++x) {
        const int w = wt + 2 + x;
        f2wait(c.w_full(w % kWStg), (uint32_t)(w / kWStg) & 1u, 3, c.oi);
    }
    tc_fence_after();
    for (int kk = 0; kk < 8; ++kk) {
        // keys [16 kk, +16): chunk kk / 2 of P (32 columns: 16 hi + 16 lo), half kk % 2 -> 8 packed columns each
        const uint32_t voff = (uint32_t)kk * 2048u;          // 16 key rows of 128 bytes
#pragma unroll
        for (int prod = 0; prod < 3; ++prod) {
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                const uint32_t acc = c.tmem_base + kTmO2 + 64u * x;
                const uint32_t vh = c.w_stage((wt + 2 + x) % kWStg) + voff, vl = vh + kPlane;
                const uint32_t ph = c.tmem_base + kTmS2 + 128u * x + 32u * (kk >> 1) + 8u * (kk & 1), pl = ph + 16u;
                if (prod == 0) umma_bf16_ta(acc, pl, make_desc(vh), idesc_o, kk ? 1u : 0u);
                else if (prod == 1) umma_bf16_ta(acc, ph, make_desc(vl), idesc_o, 1u);
                else umma_bf16_ta(acc, ph, make_desc(vh), idesc_o, 1u);
            }
        }
    }
    umma_commit(c.o_full(0));
    umma_commit(c.o_full(1));
    for (int x = 0; x < 2; ++x) umma_commit_mc(c.w_empty((wt + 2 + x) % kWStg), c.mask_w);
}

__device__ __noinline__ void attn_workers(const Ctx2& c, const F2Op* gop, const F2Fields& op, int na) {
    const int q = c.warp & 3, x = c.warp >> 2;             // quadrant q of head x
    const uint32_t par = (uint32_t)na & 1u;
    const uint32_t tm_q = c.tmem_base + ((uint32_t)(q * 32) << 16);
    const int row = 32 * q + c.lane;
    const int i = row & 63;                                // position inside the sequence
    const bool rowv = i < c.t;
    const int blk = q >> 1;                                // 64-key block that holds this row's sequence
    const int cfirst = 2 * blk;
    const int nch = (q & 1) + 1;                           // 32-key chunks with visible keys (causal)
    const int head = 2 * c.r + x;
    const float slope = __ldg(op.slopes + head);
    const uint32_t tS = tm_q + kTmS2 + 128u * x;
    f2wait(c.s_full(), par, 7, c.oi);
    tc_fence_after();
    {
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
            const float e0 = __expf(__uint_as_float(r0[e]) - mm);
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
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (c.lane == 0) mbar_arrive(c.p_ready(x));
    // ---- O of head x, rows of quadrant q -> planes (64 columns = one 128-byte plane row per thread)
    f2wait(c.o_full(x), par, 9, c.oi);
    tc_fence_after();
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        uint32_t raw[32], H[16], L[16];
        tmem_ld32(tm_q + kTmO2 + 64u * x + 32u * half, raw);
#pragma unroll
        for (int e = 0; e < 16; ++e) split2(__uint_as_float(raw[2 * e]), __uint_as_float(raw[2 * e + 1]), H[e], L[e]);
        stg_wait_read(c.lane);
        store_block_halves(c.stg(), c.lane, H, L, &gop->m[4], &gop->m[5], head * 64 + 32 * half, 32 * (q & 1), 2 * c.b + (q >> 1));
    }
    tc_fence_before();
    stores_done(c.lane);
}

// ============================== ring gather (+ downsample tail) ==============================
// Channel r of the stream: X rows 64r + j = ring rows oldest first (vap_main.py:274-283), zero rows above t.  One warp per
// row; a lane owns floats [4 lane, 4 lane + 4) and [128 + 4 lane, 128 + 4 lane + 4), so that every load and store of the
// warp is one contiguous run of full 32-byte sectors (lane-owns-8-consecutive-floats made every fp32 store a half-sector
// write): fp32 row, bf16 planes and the LayerNorm block statistics (8 lanes per 32-column block, two blocks per lane).
__device__ __forceinline__ void st_global_v2(void* p, uint32_t a, uint32_t b) { asm volatile("st.global.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void emit_row(const Fused2Params& p, size_t grow, int lane, const float4 a, const float4 b) {
    float* xf = p.Xf + grow * kD + 4 * lane;
    *reinterpret_cast<float4*>(xf) = a;
    *reinterpret_cast<float4*>(xf + 128) = b;
    uint32_t h[4], l[4];
    split2(a.x, a.y, h[0], l[0]);
    split2(a.z, a.w, h[1], l[1]);
    split2(b.x, b.y, h[2], l[2]);
    split2(b.z, b.w, h[3], l[3]);
    __nv_bfloat16* xh = p.Xh + grow * kD + 4 * lane;
    __nv_bfloat16* xl = p.Xl + grow * kD + 4 * lane;
    st_global_v2(xh, h[0], h[1]);
    st_global_v2(xh + 128, h[2], h[3]);
    st_global_v2(xl, l[0], l[1]);
    st_global_v2(xl + 128, l[2], l[3]);
    float s0 = (a.x + a.y) + (a.z + a.w), s1 = (b.x + b.y) + (b.z + b.w);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    const float m0 = s0 * (1.0f / 32.0f), m1 = s1 * (1.0f / 32.0f);
    float q0 = 0.f, q1 = 0.f;
    { float d; d = a.x - m0; q0 = fmaf(d, d, q0); d = a.y - m0; q0 = fmaf(d, d, q0); d = a.z - m0; q0 = fmaf(d, d, q0); d = a.w - m0; q0 = fmaf(d, d, q0); }
    { float d; d = b.x - m1; q1 = fmaf(d, d, q1); d = b.y - m1; q1 = fmaf(d, d, q1); d = b.z - m1; q1 = fmaf(d, d, q1); d = b.w - m1; q1 = fmaf(d, d, q1); }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        q0 += __shfl_xor_sync(0xffffffffu, q0, o);
        q1 += __shfl_xor_sync(0xffffffffu, q1, o);
    }
    if ((lane & 7) == 0) {
        *reinterpret_cast<float2*>(p.stats + grow * 16 + (lane >> 3) * 2) = make_float2(m0, q0);
        *reinterpret_cast<float2*>(p.stats + grow * 16 + (4 + (lane >> 3)) * 2) = make_float2(m1, q1);
    }
}

__device__ __noinline__ void gather_op(const Ctx2& c, const Fused2Params& p, int id, int cnt) {
    const int ch = c.r;
    const float* rg = p.ring + ((size_t)id * 2 + ch) * p.T * kD;
    const size_t grow0 = (size_t)c.b * 128 + 64 * ch;
    const bool ds = p.ds_part != nullptr && !c.ghost;
    const int jnew = ds ? c.t - 1 : -1;
    // The warp's 8 rows (warp + 8u), two at a time.  What paces this op is not the loop shape (8 rows in flight: the same
    // 24 k cycles) but the stores: 128 KB per CTA at ~2 TB/s chip-wide, the rate at which L2 can evict dirty lines to make
    // room (timing-only build without the stores: 15 k), and the tensor-map prefetch of the two idle warps (20-24 k).
#pragma unroll 1
    for (int u0 = 0; u0 < 8; u0 += 2) {
        float4 va[2], vb[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = c.warp + (u0 + u) * kWorkers2;
            va[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            vb[u] = va[u];
            if (j < c.t && j != jnew) {
                const int slot = (cnt - c.t + j) % p.T;
                va[u] = __ldg(reinterpret_cast<const float4*>(rg + (size_t)slot * kD + 4 * c.lane));
                vb[u] = __ldg(reinterpret_cast<const float4*>(rg + (size_t)slot * kD + 128 + 4 * c.lane));
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = c.warp + (u0 + u) * kWorkers2;
            if (j != jnew) emit_row(p, grow0 + j, c.lane, va[u], vb[u]);
        }
    }
    if (c.fine && c.tid == 0) c.fine[2] = clock64();                                                    // debug stamp: rows emitted
    if (c.r == 0 && c.tid == 0 && !c.ghost) p.tvalid[c.b] = c.t;
}

// The newest embedding of channel r (one warp, the side-task warp, while the workers gather the older rows): split-K sum of
// the downsample GEMM, LayerNorm, exact GELU (encoder_components.py:496-511) -> ring slot, e_out, X row t - 1.
__device__ __noinline__ void ds_tail_op(const Ctx2& c, const Fused2Params& p, int id, int cnt) {
    const int ch = c.r;
    const size_t grow0 = (size_t)c.b * 128 + 64 * ch;
    const bool ds = p.ds_part != nullptr && !c.ghost;
    if (ds) {
        const int n = 2 * c.b + ch;
        const float* pr = p.ds_part + (size_t)n * kD + 4 * c.lane;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b4 = a;
        for (int z0 = 0; z0 < p.ds_nsplit; z0 += 8) {
            float4 pa[8], pb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float* q8 = pr + (size_t)min(z0 + u, p.ds_nsplit - 1) * p.ds_stride;
                pa[u] = __ldcg(reinterpret_cast<const float4*>(q8));
                pb[u] = __ldcg(reinterpret_cast<const float4*>(q8 + 128));
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
        const float mean = warp_sum2(sum) * (1.0f / 256.0f);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float d = v8[i] - mean;
            sq = fmaf(d, d, sq);
        }
        const float rstd = 1.0f / sqrtf(warp_sum2(sq) * (1.0f / 256.0f) + 1e-5f);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.ds_lnw + 4 * c.lane)), w1 = __ldg(reinterpret_cast<const float4*>(p.ds_lnw + 128 + 4 * c.lane));
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ds_lnb + 4 * c.lane)), g1 = __ldg(reinterpret_cast<const float4*>(p.ds_lnb + 128 + 4 * c.lane));
        const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w}, bb[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) v8[i] = gelu_erf((v8[i] - mean) * rstd * ww[i] + bb[i]);
        const float4 o0 = make_float4(v8[0], v8[1], v8[2], v8[3]), o1 = make_float4(v8[4], v8[5], v8[6], v8[7]);
        float* dst[2] = {p.ring_w + (((size_t)id * 2 + ch) * p.T + (cnt - 1) % p.T) * kD, p.e_out ? p.e_out + (size_t)n * kD : nullptr};
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (dst[k]) {
                *reinterpret_cast<float4*>(dst[k] + 4 * c.lane) = o0;
                *reinterpret_cast<float4*>(dst[k] + 128 + 4 * c.lane) = o1;
            }
        emit_row(p, grow0 + (c.t - 1), c.lane, o0, o1);
    }
    if (c.fine && c.lane == 0) c.fine[3] = clock64();                                                   // debug stamp: newest frame done
}

// ============================== op loop ==============================
__device__ __forceinline__ void op_sync(int cta_only) {
#ifndef VAPB_F2_NOFENCE
    fence_async_global();
#endif
    if (cta_only) {
        asm volatile("bar.sync 1, %0;" ::"n"(kThreads2) : "memory");
    } else {
        cl_arrive2();
        cl_wait2();
    }
}

__global__ void __launch_bounds__(kThreads2, 1) k_stream_tf2(const Fused2Params p) {
    extern __shared__ uint8_t smem_raw[];
    Ctx2 c;
    c.sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    c.bars = c.sbase + (kAStages + kWStg) * kStage + kWorkers2 * kStgBytes;
    c.opslot = reinterpret_cast<F2Fields*>(smem_raw + (c.bars - smem_u32(smem_raw)) + 256);
    c.tid = threadIdx.x;
    c.warp = c.tid >> 5;
    c.lane = c.tid & 31;
    const int krank = (int)cluster_ctarank();
    c.r = krank & 1;
    c.sidx = krank >> 1;
    c.b = 2 * (int)(blockIdx.x >> 2) + c.sidx;
    c.ghost = c.b >= p.B;
    c.mask_a = (uint16_t)(3u << (2 * c.sidx));
    c.mask_w = (uint16_t)(5u << c.r);
    c.T = p.T;
    c.oi = 0;
    c.fine = nullptr;
    const int id = __ldg(p.ids + (c.ghost ? p.B - 1 : c.b));
    const int cnt = __ldg(p.count + id) + 1;           // frames including the one appended this step
    c.t = cnt < p.T ? cnt : p.T;

    if (c.warp == kWorkers2 && c.lane == 0) {
        for (int i = 0; i < kAStages; ++i) { mbar_init(c.a_full(i), 1); mbar_init(c.a_empty(i), 2); }      // "empty": both consumers of a multicast tile
        for (int i = 0; i < kWStg; ++i) { mbar_init(c.w_full(i), 1); mbar_init(c.w_empty(i), 2); }
        for (int i = 0; i < kAcc; ++i) { mbar_init(c.acc_full(i), 1); mbar_init(c.acc_empty(i), kWorkers2); }
        mbar_init(c.s_full(), 1);
        for (int i = 0; i < 2; ++i) { mbar_init(c.p_ready(i), 4); mbar_init(c.o_full(i), 1); }
        fence_barrier_init();
    }
    if (c.warp == kWorkers2 + 1) tmem_alloc(c.tmem_slot(), 512u);
    if (c.warp == 0) reinterpret_cast<uint32_t*>(c.opslot)[c.lane] = __ldg(reinterpret_cast<const uint32_t*>(&p.ops[0].f) + c.lane);
    tc_fence_before();
    cluster_sync_all();          // every CTA's barriers are initialised before any multicast copy or remote commit can reach them
    tc_fence_after();
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c.tmem_base) : "r"(c.tmem_slot()));

    const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && c.tid == 0;

#ifndef VAPB_F2_NO_TMAP_PREFETCH
    // the op program (7 tensor maps per op) is as cold as everything else when a step starts: fetch all of it now, not at
    // the first TMA instruction of each op (-7 us per step).  The MMA and A-producer warps do it: they idle through op 0
    // (the gather), and a prefetch.tensormap holds its warp for a while (issued by the side warp in front of the
    // newest-frame tail they made the gather 10 k cycles longer).
    if (c.warp == kWorkers2 + 1 || c.warp == kWorkers2 + 2)
        for (int i = (c.warp - kWorkers2 - 1) * 32 + c.lane; i < (p.n_ops - 1) * 7; i += 64) prefetch_tmap(&p.ops[1 + i / 7].m[i % 7]);      // op 0 has no maps
#endif
    if (c.warp == kWorkers2) {
        // ---------------- W producer: free-running, reads the op list from global memory ----------------
        int wt = 0;
        bool pending = false;                  // a cluster-barrier arrive of this warp is outstanding
        for (int oi = 0; oi < p.n_ops; ++oi) {
            c.oi = oi;
            const F2Fields* gf = &p.ops[oi].f;
            const int kind = __shfl_sync(0xffffffffu, __ldg(&gf->kind), 0);
            const int N = __shfl_sync(0xffffffffu, __ldg(&gf->N), 0), K = __shfl_sync(0xffffffffu, __ldg(&gf->K), 0);
            const int cta_sync = __shfl_sync(0xffffffffu, __ldg(&gf->cta_sync), 0);
            c.fine = (p.dbg != nullptr && blockIdx.x == 0 && oi == p.dbg_op) ? p.dbg + 40 : nullptr;
#ifdef VAPB_F2_TRACE
            if (blockIdx.x < 2 && c.lane == 0 && oi < 4) printf("W-warp block %d op %d kind %d N %d K %d sync %d wt %d\n", blockIdx.x, oi, kind, N, K, cta_sync, wt);
#endif
            if (kind == F2_GEMM) {
                if (elect_one()) gemm_tma_w(c, &p.ops[oi], N, K, wt);
                wt += (N >> 8) * (K >> 6);
            } else if (kind == F2_ATTN) {
                // K0 K1 V0 V1 are filled by warp 10 at these four ring positions.  This warp still waits for them to be FREE, as
                // if it filled them itself: a parity wait cannot tell "two phases behind" from "up to date", so a producer
                // that skipped the wait could wrap around a stage whose previous tile has not been consumed yet.
                for (int i = 0; i < 4; ++i) {
                    const int w = wt + i;
                    f2wait(c.w_empty(w % kWStg), ((uint32_t)(w / kWStg) & 1u) ^ 1u, 4, c.oi);
                }
                wt += 4;
            }
            __syncwarp();
            // barrier behind op oi: arrive without waiting (the wait for the previous cluster barrier keeps this warp at most
            // one cluster barrier ahead; CTA barriers are never adjacent in the op list)
            if (cta_sync) {
                asm volatile("bar.arrive 1, %0;" ::"n"(kThreads2) : "memory");      // named barrier 1 = the CTA-level op barrier
            } else {
                if (pending) cl_wait2();
                cl_arrive2();
                pending = true;
            }
        }
        if (pending) cl_wait2();
    } else {
        int at = 0, wt = 0, gs = 0, na = 0;      // running counters: A tiles, W tiles, accumulator subtiles, attentions
        for (int oi = 0; oi < p.n_ops; ++oi) {
            const F2Fields& op = c.opslot[oi & 1];
            c.oi = oi;
            const int kind = __shfl_sync(0xffffffffu, op.kind, 0);
            const int cta_sync = __shfl_sync(0xffffffffu, op.cta_sync, 0);
            if (dbg) p.dbg[oi] = clock64();
            c.fine = (p.dbg != nullptr && blockIdx.x == 0 && oi == p.dbg_op) ? p.dbg + 40 : nullptr;
            if (c.fine && c.tid == 0) c.fine[0] = clock64();
            if (c.warp < kWorkers2) {
                if (kind == F2_GEMM) gemm_epilogue(c, op, p, gs);
                else if (kind == F2_ATTN) attn_workers(c, &p.ops[oi], op, na);
                else gather_op(c, p, id, cnt);
            } else if (c.warp == kWorkers2 + 1) {
                if (oi + 1 < p.n_ops)        // fields of the next op -> the other shared-memory slot (visible after the op barrier)
                    reinterpret_cast<uint32_t*>(&c.opslot[(oi + 1) & 1])[c.lane] = __ldg(reinterpret_cast<const uint32_t*>(&p.ops[oi + 1].f) + c.lane);
                if (kind == F2_GEMM && elect_one()) gemm_mma(c, op, at, wt, gs);
                if (kind == F2_ATTN && elect_one()) attn_mma(c, at, wt, na);
                __syncwarp();
            } else if (c.warp == kWorkers2 + 2) {
                if (kind == F2_GEMM && elect_one()) gemm_tma_a(c, &p.ops[oi], op, at);
                if (kind == F2_ATTN && elect_one()) attn_tma(c, &p.ops[oi], op, at, wt);
                __syncwarp();
            } else {
                const int side = c.ghost ? F2_SIDE_NONE : __shfl_sync(0xffffffffu, op.side, 0);
                if (kind == F2_GATHER) {
                    ds_tail_op(c, p, id, cnt);
                } else if (side == F2_SIDE_VAD) {
                    // vad = sigmoid(va_classifier(x[t-1])) on the ar_channel output (vap_main.py:292-293, 313-314)
                    const float* xr = p.Xf + ((size_t)c.b * 128 + 64 * c.r + (c.t - 1)) * kD + 8 * c.lane;
                    const float4 x0 = __ldcg(reinterpret_cast<const float4*>(xr)), x1 = __ldcg(reinterpret_cast<const float4*>(xr + 4));
                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.va_w + 8 * c.lane));
                    const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.va_w + 8 * c.lane + 4));
                    float s = 0.f;
                    s = fmaf(x0.x, w0.x, s); s = fmaf(x0.y, w0.y, s); s = fmaf(x0.z, w0.z, s); s = fmaf(x0.w, w0.w, s);
                    s = fmaf(x1.x, w1.x, s); s = fmaf(x1.y, w1.y, s); s = fmaf(x1.z, w1.z, s); s = fmaf(x1.w, w1.w, s);
                    s = warp_sum2(s) + __ldg(p.va_b);
                    if (c.lane == 0) (p.io ? p.io->out : p.out)[c.b * 6 + 4 + c.r] = 1.0f / (1.0f + expf(-s));
                } else if (side == F2_SIDE_GATHER_LAST) {
                    const float* xr = p.Xf + ((size_t)c.b * 128 + 64 * c.r + (c.t - 1)) * kD + 8 * c.lane;
                    float* xo = p.Xlast + (size_t)(2 * c.b + c.r) * kD + 8 * c.lane;
                    *reinterpret_cast<float4*>(xo) = __ldcg(reinterpret_cast<const float4*>(xr));
                    *reinterpret_cast<float4*>(xo + 4) = __ldcg(reinterpret_cast<const float4*>(xr + 4));
                }
                __syncwarp();
            }
            // every role advances the running counters identically
            if (kind == F2_GEMM) {
                at += gemm_a_tiles(op);
                wt += gemm_w_tiles(op);
                gs += op.N >> 8;
            } else if (kind == F2_ATTN) {
                at += 2;
                wt += 4;
                na += 1;
            }
            op_sync(cta_sync);
            if (c.fine && c.tid == 0) c.fine[4] = clock64();            // barrier passed
        }
        if (dbg) p.dbg[p.n_ops] = clock64();
    }
    tc_fence_before();
    cluster_sync_all();          // no CTA leaves while a peer's commits or copies may still target it
    if (c.warp == kWorkers2 + 1) tmem_dealloc(c.tmem_base, 512u);
}

}  // namespace

size_t fused2_smem_bytes() { return kSmem2; }

bool fused2_prepare(std::string& err) {
    static OncePerDevice once;
    if (!once.first()) return true;
    cudaError_t e = cudaFuncSetAttribute(k_stream_tf2, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2);
    if (e != cudaSuccess) {
        err = std::string("cudaFuncSetAttribute(k_stream_tf2) failed: ") + cudaGetErrorString(e);
        return false;
    }
    return true;
}

cudaError_t launch_fused_tf2(const Fused2Params& p, int B, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * ((B + 1) / 2));
    cfg.blockDim = dim3(kThreads2);
    cfg.dynamicSmemBytes = kSmem2;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_stream_tf2, p);
}

}  // namespace vapb
