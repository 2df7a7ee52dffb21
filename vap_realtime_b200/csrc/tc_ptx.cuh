// PTX wrappers shared by the tcgen05 kernels (mbarrier, TMA, TMEM, UMMA descriptors, bf16 hi/lo split).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>

namespace vapb {
namespace tcp {

constexpr int kBM = 128;          // rows per CTA tile = UMMA M
constexpr int kBK = 64;           // K per stage: 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;        // K per tcgen05.mma for 16-bit inputs
constexpr int kATile = kBM * kBK * 2;     // bytes per A plane per 64-wide k-block (16 KB)

// ---- PTX wrappers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// VAPB_WAIT_POLLS: failed polls before a wait gives up (0 = wait forever; compute-sanitizer builds use 0 because the
// instrumented producer warps are orders of magnitude slower than the polling warp).
#ifndef VAPB_WAIT_POLLS
#define VAPB_WAIT_POLLS 40000000u
#endif
// Bounded wait: a mis-programmed pipeline must fail loudly (trap -> launch failure), never hang the GPU.
// The loop counts polls instead of reading the clock (CS2R shares the XU pipe with the bf16 conversions of the
// producer warps) and the failure path is a noreturn call, so nothing is live across the printf.
static __device__ __noinline__ __attribute__((noreturn)) void mbar_wait_fail() {
    printf("vapb gemm_tc: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
    __trap();
    for (;;) {}
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (VAPB_WAIT_POLLS != 0u && ++polls > VAPB_WAIT_POLLS) mbar_wait_fail();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// TMA STORE: one box from (swizzled) shared memory to the tensor; bulk-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0),
                 "r"(c1), "r"(src)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(src)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources may be reused
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }         // writes are complete

// Multicast variant: the box lands at the same smem offset in every CTA of `mask` and completes tx bytes on
// the mbarrier at the same offset in each of them.
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar), "h"(mask)
        : "memory");
}
// tcgen05.commit that arrives on the mbarrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same with the A operand in TENSOR MEMORY (lane = row, each 32-bit column holds two consecutive-K bf16,
// 8 columns per K = 16 step): no shared-memory read for A at all.
__device__ __forceinline__ void umma_bf16_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// One elected lane of a converged warp.  ptxas knows that code guarded by elect.sync runs in a single thread, so
// the tcgen05 / TMA operands computed inside need no uniformity proof (see the note below).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// NOTE on issuing tcgen05.mma / TMA from one lane: guard the region with elect_one(), not with `lane == 0`.  Inside a
// plain divergent branch ptxas cannot prove the descriptors warp-uniform and wraps EVERY tcgen05.mma in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop (~94 cycles per MMA whatever its width, measured).
// 32 lanes x 16 columns: thread l of warp w writes v[0..15] to TMEM lane 32 * (w % 4) + l, columns taddr .. taddr + 15
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 8 columns, without the wait: the caller overlaps the load with arithmetic and calls tmem_ld_wait() before using v
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (= 1, unused for swizzled K-major)
//   [32,46) SBO >> 4 = 1024 B between 8-row groups | [46,48) version = 1 | [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = BN.
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Exact-erf GELU of a block of values, erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, the level of erff itself).
// Written stage by stage over groups of 8 so that eight independent dependency chains are in flight: the epilogue
// warps are few (2 per scheduler) and a GELU is a ~15-instruction chain through two MUFU ops, so without explicit
// instruction-level parallelism the FFN1 epilogue was latency-bound (measured: 31 k of the 49 k cycles of that op).
__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <int N>
__device__ __forceinline__ void gelu_block(float (&v)[N]) {
    static_assert(N % 8 == 0, "gelu_block works on groups of 8");
#pragma unroll
    for (int b = 0; b < N; b += 8) {
        float z[8], t[8], e[8], pl[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = fabsf(v[b + i]) * 0.70710678118654752440f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = rcp_fast(fmaf(0.3275911f, z[i], 1.0f));
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = ex2_fast(z[i] * z[i] * -1.4426950408889634f);        // exp(-z^2)
#pragma unroll
        for (int i = 0; i < 8; ++i) pl[i] = fmaf(1.061405429f, t[i], -1.453152027f);
#pragma unroll
        for (int i = 0; i < 8; ++i) pl[i] = fmaf(pl[i], t[i], 1.421413741f);
#pragma unroll
        for (int i = 0; i < 8; ++i) pl[i] = fmaf(pl[i], t[i], -0.284496736f);
#pragma unroll
        for (int i = 0; i < 8; ++i) pl[i] = fmaf(pl[i], t[i], 0.254829592f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float er = fmaf(-pl[i] * t[i], e[i], 1.0f);                  // erf(|x| / sqrt 2)
            v[b + i] = fmaf(0.5f * fabsf(v[b + i]), er, 0.5f * v[b + i]);      // 0.5 x (1 + sign(x) erf(|x| / sqrt 2))
        }
    }
}

// x = hi + lo with hi = bf16_rn(x), lo = bf16_rn(x - hi); two elements per packed conversion
// (cvt.rn.bf16x2.f32 d, a, b puts a in the upper and b in the lower half of d).
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    uint32_t h, l;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
    const float h0 = __uint_as_float(h << 16), h1 = __uint_as_float(h & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(x1 - h1), "f"(x0 - h0));
    hi = h;
    lo = l;
}


}  // namespace tcp
}  // namespace vapb
