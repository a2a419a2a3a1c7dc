// Library identification for the C ABI (include/densefusion_b200.h).
#include "df_common.cuh"
#include "../../include/densefusion_b200.h"

#ifndef DF_HAVE_TC
#define DF_HAVE_TC 0
#endif

extern "C" int df_abi_version(void) { return 1; }
extern "C" int df_features(void) { return DF_HAVE_TC ? 1 : 0; }

// ---- measured-peak probe (bench.py: roofline denominator of the fp32-issue-bound kernels K3 / K4) ----------------------
// Every thread runs 8 independent FFMA chains of `iters` steps: 16 * iters flops per thread, no memory traffic.
namespace {
__global__ void __launch_bounds__(256) ffma_probe_kernel(float* __restrict__ sink, int iters)
{
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (float)(threadIdx.x + i) * 1e-3f;
    const float m = 1.0000001f, c = 1e-7f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], m, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678f) sink[0] = s;            // never true: keeps the chains alive
}
}  // namespace

extern "C" long long df_probe_ffma(float* sink, int blocks, int iters, void* stream)
{
    if (!sink || blocks <= 0 || iters <= 0) return DF_ERR_ARG;
    ffma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(sink, iters);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return -(long long)e - 1000;
    return (long long)blocks * 256 * 16 * iters;  // flops launched
}
