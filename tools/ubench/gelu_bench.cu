// Micro-benchmark: cycles per GELU under the stream kernels' epilogue conditions (8 warps per SM = 2 per scheduler,
// 32 values per thread held in registers).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a gelu_bench.cu -o gelu_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "../../vap_realtime_b200/csrc/tc_ptx.cuh"
using namespace vapb::tcp;

__device__ __forceinline__ float gelu_as(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
    float pl = fmaf(1.061405429f, t, -1.453152027f);
    pl = fmaf(pl, t, 1.421413741f);
    pl = fmaf(pl, t, -0.284496736f);
    pl = fmaf(pl, t, 0.254829592f);
    const float er = 1.0f - pl * t * __expf(-z * z);
    return 0.5f * x + 0.5f * fabsf(x) * er;
}
template <int MODE>
__global__ void k(const float* in, float* out, long long* clk, int iters) {
    float v[32];
    for (int e = 0; e < 32; ++e) v[e] = in[(blockIdx.x * blockDim.x + threadIdx.x) * 32 + e];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = gelu_as(v[e]) + 0.25f;
        } else if (MODE == 1) {
            gelu_block(v);
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] += 0.25f;
        } else if (MODE == 2) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = gelu_erf(v[e]) + 0.25f;
        } else {            // arithmetic floor: 16 dependent FMAs per value, no MUFU
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                float a = v[e];
#pragma unroll
                for (int j = 0; j < 16; ++j) a = fmaf(a, 0.999f, 0.001f);
                v[e] = a;
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int e = 0; e < 32; ++e) s += v[e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[MODE] = t1 - t0;
}
int main() {
    const int threads = 256, blocks = 148, iters = 6;
    float *in, *out; long long* clk;
    cudaMalloc(&in, blocks * threads * 32 * 4); cudaMalloc(&out, blocks * threads * 4); cudaMallocManaged(&clk, 64);
    cudaMemset(in, 0, blocks * threads * 32 * 4);
    for (int rep = 0; rep < 2; ++rep) {
        k<0><<<blocks, threads>>>(in, out, clk, iters); k<1><<<blocks, threads>>>(in, out, clk, iters);
        k<2><<<blocks, threads>>>(in, out, clk, iters); k<3><<<blocks, threads>>>(in, out, clk, iters);
        cudaDeviceSynchronize();
    }
    const char* names[4] = {"gelu_as (v1)", "gelu_block (8 chains)", "erff", "16 dependent FMAs"};
    for (int m = 0; m < 4; ++m) printf("%-24s %8.1f cycles per 32 values per thread (8 warps/SM) = %.1f per value\n", names[m], (double)clk[m] / iters, (double)clk[m] / iters / 32);
    return 0;
}
