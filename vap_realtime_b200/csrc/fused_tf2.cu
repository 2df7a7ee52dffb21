// Per-stream persistent transformer kernel, second generation -- see fused_tf2.cuh for the scheme.
//
// CTA anatomy (384 threads, 1 CTA per SM, cluster of 2 CTAs per stream):
//   warps 0-7  workers: GEMM epilogues (tcgen05.ld -> LayerNorm correction / GELU / residual -> bf16 hi / lo planes,
//              thread-per-row, straight to global), softmax, ring gather
//   warp 8     TMA producer: A / Q / K tiles into the 4 "A" stages, W / V tiles into the 3-stage W ring
//   warp 9     TMEM owner + the single thread that issues tcgen05.mma
//   warp 10    side tasks (vad, newest-frame gather); warp 11 completes the warpgroup
// Every stage holds the hi and the lo plane of one 128 x 64 bf16 tile (2 x 16 KB, 128 B swizzle).
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
constexpr int kAStages = 4, kWStg = 3, kAcc = 4;      // 4 accumulators of 128 columns = all of tensor memory
constexpr int kSmem2 = (kAStages + kWStg) * kStage + 1024 /*alignment slack*/ + 512 /*barriers, op slots*/;
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
    int b, r;              // stream index in the batch, CTA rank in the cluster
    int T, t, oi;
    __device__ __forceinline__ uint32_t a_stage(int s) const { return sbase + (uint32_t)s * kStage; }
    __device__ __forceinline__ uint32_t w_stage(int s) const { return sbase + (uint32_t)(kAStages + s) * kStage; }
    __device__ __forceinline__ uint32_t a_full(int s) const { return bars + 8u * s; }
    __device__ __forceinline__ uint32_t a_empty(int s) const { return bars + 32u + 8u * s; }
    __device__ __forceinline__ uint32_t w_full(int s) const { return bars + 64u + 8u * s; }
    __device__ __forceinline__ uint32_t w_empty(int s) const { return bars + 88u + 8u * s; }
    __device__ __forceinline__ uint32_t acc_full(int s) const { return bars + 112u + 8u * s; }
    __device__ __forceinline__ uint32_t acc_empty(int s) const { return bars + 144u + 8u * s; }
    __device__ __forceinline__ uint32_t s_full() const { return bars + 176u; }
    __device__ __forceinline__ uint32_t p_ready(int x) const { return bars + 184u + 8u * x; }
    __device__ __forceinline__ uint32_t o_full(int x) const { return bars + 200u + 8u * x; }
    __device__ __forceinline__ uint32_t tmem_slot() const { return bars + 216u; }
};

// exact-erf GELU, erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7), as in fused_tf.cu
__device__ __forceinline__ float gelu_as2(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
    float pl = fmaf(1.061405429f, t, -1.453152027f);
    pl = fmaf(pl, t, 1.421413741f);
    pl = fmaf(pl, t, -0.284496736f);
    pl = fmaf(pl, t, 0.254829592f);
    const float er = 1.0f - pl * t * __expf(-z * z);
    return 0.5f * x + 0.5f * fabsf(x) * er;
}

__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// 256-bit global accesses (sm_100): a thread-per-row access touches one 128 B line per lane, and the LSU pays per line
// visited, so fewer, wider instructions are what counts here
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* a) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]),
                 "r"(a[5]), "r"(a[6]), "r"(a[7])
                 : "memory");
}
__device__ __forceinline__ void st_global_v8f(float* p, const float* a) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(a[4]),
                 "f"(a[5]), "f"(a[6]), "f"(a[7])
                 : "memory");
}
__device__ __forceinline__ void ld_global_cg_v8f(const float* p, float* a) {
    asm volatile("ld.global.cg.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7])
                 : "l"(p));
}
#ifdef VAPB_F2_NOSTORE        // timing experiment only: results are wrong
#define F2_STORE(x)
#else
#define F2_STORE(x) x
#endif

// 32 consecutive fp32 values of one row -> 64 bytes in the hi plane and 64 bytes in the lo plane
__device__ __forceinline__ void store_planes32(const float (&v)[32], __nv_bfloat16* hi, __nv_bfloat16* lo) {
    uint32_t h[16], l[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) split2(v[2 * e], v[2 * e + 1], h[e], l[e]);
#ifdef VAPB_F2_V4
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
        F2_STORE(st_global_v4(hi + 8 * q4, h[4 * q4], h[4 * q4 + 1], h[4 * q4 + 2], h[4 * q4 + 3]);)
        F2_STORE(st_global_v4(lo + 8 * q4, l[4 * q4], l[4 * q4 + 1], l[4 * q4 + 2], l[4 * q4 + 3]);)
    }
#else
    F2_STORE(st_global_v8(hi, h);)
    F2_STORE(st_global_v8(hi + 16, h + 8);)
    F2_STORE(st_global_v8(lo, l);)
    F2_STORE(st_global_v8(lo + 16, l + 8);)
#endif
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
__device__ __forceinline__ void gemm_epilogue(const Ctx2& c, const F2Fields& op, const Fused2Params& p, int gs) {
    const int q = c.warp & 3, hf = c.warp >> 2;
    const int row = 32 * q + c.lane;                       // tile row = TMEM lane
    const size_t grow = (size_t)c.b * 128 + row;
    const int ns = op.N >> 8;
    const uint32_t tm_row = c.tmem_base + ((uint32_t)(q * 32) << 16);
    float mu = 0.f, rstd = 1.f;
    if (op.n_ln > 0) row_stats(p.stats + grow * 16, mu, rstd);
    const int seq = row >> 6, pos = row & 63;
    long long* fine = c.tid == 0 ? c.fine : nullptr;
    for (int s = 0; s < ns; ++s) {
        const int g = gs + s, slot = g & (kAcc - 1);
        const int n0 = 256 * s + 128 * c.r + 64 * hf;
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
            const int col = n0 + 32 * blk;
            float v[32];
            float xr[32];
            if (op.out_mode == F2_OUT_X) {                 // residual rows requested before the accumulator is waited for
#pragma unroll
                for (int e = 0; e < 4; ++e) ld_global_cg_v8f(p.Xf + grow * kD + col + 8 * e, xr + 8 * e);
            }
            if (blk == 0) {
                f2wait(c.acc_full(slot), (uint32_t)(g / kAcc) & 1u, 5, c.oi);
                tc_fence_after();
                if (fine && s == 0) fine[1] = clock64();            // first accumulator complete
                if (fine && s == ns - 1) fine[2] = clock64();       // last accumulator complete
            }
            {
                uint32_t raw[32];
                tmem_ld32(tm_row + (uint32_t)(slot * 128 + 64 * hf + 32 * blk), raw);
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(raw[e]);
            }
            if (blk == 1) {
                tc_fence_before();
                __syncwarp();
                if (c.lane == 0) mbar_arrive(c.acc_empty(slot));
            }
            if (col < op.n_ln) {
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
            if (op.act == 1) {
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = gelu_as2(v[e]);
            }
            if (op.out_mode == F2_OUT_PLANES) {
                store_planes32(v, op.out_hi + grow * op.ld_out + col, op.out_lo + grow * op.ld_out + col);
            } else if (op.out_mode == F2_OUT_X) {
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] += xr[e];
#pragma unroll
                for (int e = 0; e < 4; ++e) { F2_STORE(st_global_v8f(p.Xf + grow * kD + col + 8 * e, v + 8 * e);) }
                store_planes32(v, p.Xh + grow * kD + col, p.Xl + grow * kD + col);
                const float2 st = block_stats32(v);
                *reinterpret_cast<float2*>(p.stats + grow * 16 + (col >> 5) * 2) = st;
            } else {                                       // fp32 rows in the batched kernels' layout (tail of the pruned layer)
                if (pos < c.T) {
                    float* dst = (col < 512 ? op.out_f : op.out_f2) + ((size_t)(2 * c.b + seq) * c.T + pos) * 512 + (col & 511);
#pragma unroll
                    for (int e = 0; e < 4; ++e) { F2_STORE(st_global_v8f(dst + 8 * e, v + 8 * e);) }
                }
            }
        }
    }
    if (fine) fine[3] = clock64();                                  // epilogue of warp 0 done
}

// ============================== GEMM: TMA side ==============================
// Tile order = MMA order.  Subtiles go in groups: a PAIR of subtiles (two W tiles per k-block, two 128-column
// accumulators issued interleaved) or a SINGLE subtile (one W tile per k-block, issued as two N = 64 halves).
// K = 256: the four A tiles are loaded once and stay resident for every group; K = 768: A tiles stream with W.
__device__ __forceinline__ void gemm_tma(const Ctx2& c, const F2Op* gop, const F2Fields& op, int at, int wt) {
    const int ns = op.N >> 8, nkb = op.K >> 6;
    const bool resident = nkb <= kAStages;
    const int row0 = c.b * 128;
    int w = wt;
    if (c.fine) c.fine[7] = clock64();                              // TMA: first issue
    for (int s = 0; s < ns; s += 2) {
        const int gsz = (s + 1 < ns) ? 2 : 1;
        for (int kb = 0; kb < nkb; ++kb) {
            if (s == 0 || !resident) {
                const int ai = at + (resident ? kb : (s / 2) * nkb + kb);
                const int st = ai % kAStages;
                f2wait(c.a_empty(st), ((uint32_t)(ai / kAStages) & 1u) ^ 1u, 2, c.oi);
                mbar_arrive_expect_tx(c.a_full(st), kStage);
                tma_load_2d(c.a_stage(st), &gop->m[0], kb * kBK, row0, c.a_full(st));
                tma_load_2d(c.a_stage(st) + kPlane, &gop->m[1], kb * kBK, row0, c.a_full(st));
            }
            for (int u = 0; u < gsz; ++u, ++w) {
                const int st = w % kWStg;
                f2wait(c.w_empty(st), ((uint32_t)(w / kWStg) & 1u) ^ 1u, 4, c.oi);
                mbar_arrive_expect_tx(c.w_full(st), kStage);
                const int n0 = 256 * (s + u) + 128 * c.r;
                tma_load_2d(c.w_stage(st), &gop->m[2], kb * kBK, n0, c.w_full(st));
                tma_load_2d(c.w_stage(st) + kPlane, &gop->m[3], kb * kBK, n0, c.w_full(st));
            }
        }
    }
    if (c.fine) c.fine[8] = clock64();                              // TMA: last issue
}
__device__ __forceinline__ int gemm_a_tiles(const F2Fields& op) {
    const int ns = op.N >> 8, nkb = op.K >> 6;
    return nkb <= kAStages ? kAStages : nkb * ((ns + 1) / 2);       // always a multiple of kAStages (K = 768: 12 per group)
}
__device__ __forceinline__ int gemm_w_tiles(const F2Fields& op) { return (op.N >> 8) * (op.K >> 6); }

// ============================== GEMM: MMA side (one elected thread) ==============================
template <bool PAIR>
__device__ __forceinline__ void mma_group(const Ctx2& c, int nkb, bool resident, bool first_group, bool last_group, int at_group, int& w,
                                          int slot0, int slot1) {
    constexpr uint32_t idesc = PAIR ? make_idesc(128) : make_idesc(64);
    const uint32_t acc0 = c.tmem_base + (uint32_t)(slot0 * 128);
    const uint32_t acc1 = PAIR ? c.tmem_base + (uint32_t)(slot1 * 128) : acc0 + 64u;
    for (int kb = 0; kb < nkb; ++kb) {
        const int ai = at_group + kb;
        const int ast = ai % kAStages;
        if (first_group || !resident) {
            f2wait(c.a_full(ast), (uint32_t)(ai / kAStages) & 1u, 1, c.oi);
        }
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
            umma_bf16(acc1, make_desc(ab + kPlane + koff), make_desc(wb1 + koff), idesc, first);
            umma_bf16(acc0, make_desc(ab + koff), make_desc(wb0 + kPlane + koff), idesc, 1u);
            umma_bf16(acc1, make_desc(ab + koff), make_desc(wb1 + kPlane + koff), idesc, 1u);
            umma_bf16(acc0, make_desc(ab + koff), make_desc(wb0 + koff), idesc, 1u);
            umma_bf16(acc1, make_desc(ab + koff), make_desc(wb1 + koff), idesc, 1u);
        }
        umma_commit(c.w_empty(ws0));
        if (PAIR) umma_commit(c.w_empty(ws1));
        if (last_group || !resident) umma_commit(c.a_empty(ast));
        w += PAIR ? 2 : 1;
    }
    umma_commit(c.acc_full(slot0));
    if (PAIR) umma_commit(c.acc_full(slot1));
}

__device__ __forceinline__ void gemm_mma(const Ctx2& c, const F2Fields& op, int at, int wt, int gs) {
    const int ns = op.N >> 8, nkb = op.K >> 6;
    const bool resident = nkb <= kAStages;
    int w = wt;
    for (int s = 0; s < ns; s += 2) {
        const bool pair = s + 1 < ns;
        const int g0 = gs + s, g1 = gs + s + 1;
        const int slot0 = g0 & (kAcc - 1), slot1 = g1 & (kAcc - 1);
        f2wait(c.acc_empty(slot0), ((uint32_t)(g0 / kAcc) & 1u) ^ 1u, 6, c.oi);
        if (pair) f2wait(c.acc_empty(slot1), ((uint32_t)(g1 / kAcc) & 1u) ^ 1u, 6, c.oi);
        tc_fence_after();
        const int at_group = at + (resident ? 0 : (s / 2) * nkb);
        const bool last = s + 2 >= ns;
        if (pair) mma_group<true>(c, nkb, resident, s == 0, last, at_group, w, slot0, slot1);
        else mma_group<false>(c, nkb, resident, s == 0, last, at_group, w, slot0, slot1);
    }
    if (c.fine) c.fine[6] = clock64();                              // MMA: last issue
}

// ============================== attention (modules.py:82-110, 170-212) ==============================
// Two heads (2r, 2r + 1) per CTA.  Shared memory: Q of head x in A stage x, K of head x in A stage 2 + x, V of head x
// in the next two W ring stages.  K / V tiles hold the keys of sequence c (or of its sibling for the cross attention)
// in rows 64c .. 64c + 63, so row (c, i) of Q always finds its keys in columns 64c + j of S.
// Tensor memory: S_x / P_x at columns 128x, O_x at 256 + 64x.
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
    for (int x = 0; x < 2; ++x) {                          // K: two boxes of 64 rows (swapped for the sibling channel)
        const int ai = at + 2 + x, st = ai % kAStages;
        f2wait(c.a_empty(st), ((uint32_t)(ai / kAStages) & 1u) ^ 1u, 2, c.oi);
        mbar_arrive_expect_tx(c.a_full(st), kStage);
        const int col = op.kcol + (2 * c.r + x) * 64;
        for (int half = 0; half < 2; ++half) {
            const int src = row0 + 64 * (half ^ op.sibling);
            tma_load_2d(c.a_stage(st) + half * (kPlane / 2), &gop->m[2], col, src, c.a_full(st));
            tma_load_2d(c.a_stage(st) + kPlane + half * (kPlane / 2), &gop->m[3], col, src, c.a_full(st));
        }
    }
    for (int x = 0; x < 2; ++x) {                          // V
        const int w = wt + x, st = w % kWStg;
        f2wait(c.w_empty(st), ((uint32_t)(w / kWStg) & 1u) ^ 1u, 4, c.oi);
        mbar_arrive_expect_tx(c.w_full(st), kStage);
        const int col = op.vcol + (2 * c.r + x) * 64;
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
    for (int i = 0; i < 4; ++i) {
        const int ai = at + i;
        f2wait(c.a_full(ai % kAStages), (uint32_t)(ai / kAStages) & 1u, 1, c.oi);
    }
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t koff = (uint32_t)k * kUmmaK * 2;
#pragma unroll
        for (int prod = 0; prod < 3; ++prod) {
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                const uint32_t qb = c.a_stage((at + x) % kAStages) + koff, kb = c.a_stage((at + 2 + x) % kAStages) + koff;
                const uint32_t acc = c.tmem_base + kTmS2 + 128u * x;
                if (prod == 0) umma_bf16(acc, make_desc(qb + kPlane), make_desc(kb), idesc_s, k ? 1u : 0u);
                else if (prod == 1) umma_bf16(acc, make_desc(qb), make_desc(kb + kPlane), idesc_s, 1u);
                else umma_bf16(acc, make_desc(qb), make_desc(kb), idesc_s, 1u);
            }
        }
    }
    umma_commit(c.s_full());
    for (int i = 0; i < 4; ++i) umma_commit(c.a_empty((at + i) % kAStages));       // Q and K are free once S is complete
    f2wait(c.p_ready(0), par, 8, c.oi);
    f2wait(c.p_ready(1), par, 8, c.oi);
    for (int x = 0; x < 2; ++x) {
        const int w = wt + x;
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
                const uint32_t vh = c.w_stage((wt + x) % kWStg) + voff, vl = vh + kPlane;
                const uint32_t ph = c.tmem_base + kTmS2 + 128u * x + 32u * (kk >> 1) + 8u * (kk & 1), pl = ph + 16u;
                if (prod == 0) umma_bf16_ta(acc, pl, make_desc(vh), idesc_o, kk ? 1u : 0u);
                else if (prod == 1) umma_bf16_ta(acc, ph, make_desc(vl), idesc_o, 1u);
                else umma_bf16_ta(acc, ph, make_desc(vh), idesc_o, 1u);
            }
        }
    }
    umma_commit(c.o_full(0));
    umma_commit(c.o_full(1));
    for (int x = 0; x < 2; ++x) umma_commit(c.w_empty((wt + x) % kWStg));
}

__device__ __forceinline__ void attn_workers(const Ctx2& c, const F2Fields& op, int na) {
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
    // ---- O of head x, rows of quadrant q -> planes
    f2wait(c.o_full(x), par, 9, c.oi);
    tc_fence_after();
    const size_t grow = (size_t)c.b * 128 + row;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        uint32_t raw[32];
        tmem_ld32(tm_q + kTmO2 + 64u * x + 32u * half, raw);
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(raw[e]);
        const int col = head * 64 + 32 * half;
        store_planes32(v, op.out_hi + grow * op.ld_out + col, op.out_lo + grow * op.ld_out + col);
    }
    tc_fence_before();
}

// ============================== ring gather (+ downsample tail) ==============================
// Channel r of the stream: X rows 64r + j = ring rows oldest first (vap_main.py:274-283), zero rows above t; writes the
// fp32 rows, their bf16 planes and the LayerNorm block statistics.  Warp 7 finishes the newest embedding first
// (split-K sum of the downsample GEMM, LayerNorm, GELU: encoder_components.py:496-511) and appends it to the ring.
__device__ __forceinline__ void gather_op(const Ctx2& c, const Fused2Params& p, int id, int cnt) {
    const int ch = c.r;
    const float* rg = p.ring + ((size_t)id * 2 + ch) * p.T * kD;
    const size_t grow0 = (size_t)c.b * 128 + 64 * ch;
    const int jnew = p.ds_part ? c.t - 1 : -1;
    for (int item = c.tid; item < 64 * 8; item += kWorkers2 * 32) {
        const int j = item >> 3, blk = item & 7;
        if (j == jnew) continue;
        float v[32];
        if (j < c.t) {
            const int slot = (cnt - c.t + j) % p.T;
            const float4* src = reinterpret_cast<const float4*>(rg + (size_t)slot * kD + 32 * blk);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float4 x4 = __ldg(src + e);
                v[4 * e] = x4.x; v[4 * e + 1] = x4.y; v[4 * e + 2] = x4.z; v[4 * e + 3] = x4.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = 0.f;
        }
        const size_t grow = grow0 + j;
        float4* xo = reinterpret_cast<float4*>(p.Xf + grow * kD + 32 * blk);
#pragma unroll
        for (int e = 0; e < 8; ++e) xo[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
        store_planes32(v, p.Xh + grow * kD + 32 * blk, p.Xl + grow * kD + 32 * blk);
        *reinterpret_cast<float2*>(p.stats + grow * 16 + blk * 2) = block_stats32(v);
    }
    if (p.ds_part && c.warp == kWorkers2 - 1) {
        const int n = 2 * c.b + ch;
        const float* pr = p.ds_part + (size_t)n * kD + 8 * c.lane;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b4 = a;
        for (int z0 = 0; z0 < p.ds_nsplit; z0 += 8) {
            float4 pa[8], pb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float* q8 = pr + (size_t)min(z0 + u, p.ds_nsplit - 1) * p.ds_stride;
                pa[u] = __ldcg(reinterpret_cast<const float4*>(q8));
                pb[u] = __ldcg(reinterpret_cast<const float4*>(q8 + 4));
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
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.ds_lnw + 8 * c.lane)), w1 = __ldg(reinterpret_cast<const float4*>(p.ds_lnw + 8 * c.lane + 4));
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ds_lnb + 8 * c.lane)), g1 = __ldg(reinterpret_cast<const float4*>(p.ds_lnb + 8 * c.lane + 4));
        const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w}, bb[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) v8[i] = gelu_erf((v8[i] - mean) * rstd * ww[i] + bb[i]);
        const float4 o0 = make_float4(v8[0], v8[1], v8[2], v8[3]), o1 = make_float4(v8[4], v8[5], v8[6], v8[7]);
        const size_t grow = grow0 + (c.t - 1);
        float* dst[3] = {p.ring_w + (((size_t)id * 2 + ch) * p.T + (cnt - 1) % p.T) * kD, p.Xf + grow * kD, p.e_out ? p.e_out + (size_t)n * kD : nullptr};
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (dst[k]) {
                *reinterpret_cast<float4*>(dst[k] + 8 * c.lane) = o0;
                *reinterpret_cast<float4*>(dst[k] + 8 * c.lane + 4) = o1;
            }
        // planes: 8 values = 16 bytes per plane; block statistics: 4 lanes share one 32-column block
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split2(v8[2 * e], v8[2 * e + 1], h[e], l[e]);
        st_global_v4(p.Xh + grow * kD + 8 * c.lane, h[0], h[1], h[2], h[3]);
        st_global_v4(p.Xl + grow * kD + 8 * c.lane, l[0], l[1], l[2], l[3]);
        float bs = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) bs += v8[i];
        bs += __shfl_xor_sync(0xffffffffu, bs, 1);
        bs += __shfl_xor_sync(0xffffffffu, bs, 2);
        const float bm = bs * (1.0f / 32.0f);
        float bq = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float d = v8[i] - bm;
            bq = fmaf(d, d, bq);
        }
        bq += __shfl_xor_sync(0xffffffffu, bq, 1);
        bq += __shfl_xor_sync(0xffffffffu, bq, 2);
        if ((c.lane & 3) == 0) *reinterpret_cast<float2*>(p.stats + grow * 16 + (c.lane >> 2) * 2) = make_float2(bm, bq);
    }
    if (c.r == 0 && c.tid == 0) p.tvalid[c.b] = c.t;
}

// ============================== op loop, shared by every role ==============================
__device__ __forceinline__ void op_sync(int cta_only) {
    fence_async_global();
    if (cta_only) {
        __syncthreads();
    } else {
        cl_arrive2();
        cl_wait2();
    }
}

// (No setmaxnreg here: the op loop is shared by every role and the epilogue fits the 168-register launch bound.)
__global__ void __launch_bounds__(kThreads2, 1) k_stream_tf2(const Fused2Params p) {
    extern __shared__ uint8_t smem_raw[];
    Ctx2 c;
    c.sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    c.bars = c.sbase + (kAStages + kWStg) * kStage;
    c.opslot = reinterpret_cast<F2Fields*>(smem_raw + (c.bars - smem_u32(smem_raw)) + 256);
    c.tid = threadIdx.x;
    c.warp = c.tid >> 5;
    c.lane = c.tid & 31;
    c.r = (int)cluster_ctarank();
    c.b = blockIdx.x >> 1;
    c.T = p.T;
    c.oi = 0;
    c.fine = nullptr;
    const int id = __ldg(p.ids + c.b);
    const int cnt = __ldg(p.count + id) + 1;           // frames including the one appended this step
    c.t = cnt < p.T ? cnt : p.T;

    if (c.warp == kWorkers2 && c.lane == 0) {
        for (int i = 0; i < kAStages; ++i) { mbar_init(c.a_full(i), 1); mbar_init(c.a_empty(i), 1); }
        for (int i = 0; i < kWStg; ++i) { mbar_init(c.w_full(i), 1); mbar_init(c.w_empty(i), 1); }
        for (int i = 0; i < kAcc; ++i) { mbar_init(c.acc_full(i), 1); mbar_init(c.acc_empty(i), kWorkers2); }
        mbar_init(c.s_full(), 1);
        for (int i = 0; i < 2; ++i) { mbar_init(c.p_ready(i), 4); mbar_init(c.o_full(i), 1); }
        fence_barrier_init();
    }
    if (c.warp == kWorkers2 + 1) tmem_alloc(c.tmem_slot(), 512u);
    if (c.warp == 0) reinterpret_cast<uint32_t*>(c.opslot)[c.lane] = __ldg(reinterpret_cast<const uint32_t*>(&p.ops[0].f) + c.lane);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c.tmem_base) : "r"(c.tmem_slot()));

    const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && c.tid == 0;

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
            else if (kind == F2_ATTN) attn_workers(c, op, na);
            else gather_op(c, p, id, cnt);
        } else if (c.warp == kWorkers2) {
            if (kind == F2_GEMM && elect_one()) gemm_tma(c, &p.ops[oi], op, at, wt);
            if (kind == F2_ATTN && elect_one()) attn_tma(c, &p.ops[oi], op, at, wt);
            __syncwarp();
        } else if (c.warp == kWorkers2 + 1) {
            if (oi + 1 < p.n_ops)        // fields of the next op -> the other shared-memory slot (visible after the op barrier)
                reinterpret_cast<uint32_t*>(&c.opslot[(oi + 1) & 1])[c.lane] = __ldg(reinterpret_cast<const uint32_t*>(&p.ops[oi + 1].f) + c.lane);
            if (kind == F2_GEMM && elect_one()) gemm_mma(c, op, at, wt, gs);
            if (kind == F2_ATTN && elect_one()) attn_mma(c, at, wt, na);
            __syncwarp();
        } else if (c.warp == kWorkers2 + 2) {
            const int side = __shfl_sync(0xffffffffu, op.side, 0);
            if (side == F2_SIDE_VAD) {
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
            at += 4;
            wt += 2;
            na += 1;
        }
        op_sync(cta_sync);
        if (c.fine && c.tid == 0) c.fine[4] = clock64();            // barrier passed
    }
    if (dbg) p.dbg[p.n_ops] = clock64();
    tc_fence_before();
    __syncthreads();
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
    cfg.gridDim = dim3(2 * B);
    cfg.blockDim = dim3(kThreads2);
    cfg.dynamicSmemBytes = kSmem2;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_stream_tf2, p);
}

}  // namespace vapb
