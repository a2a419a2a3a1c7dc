// Shared helpers for the densefusion_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DF_OK 0
#define DF_ERR_ARG (-1)          // bad argument (null pointer, size out of range, ...)
#define DF_ERR_UNSUPPORTED (-2)  // valid request this build does not implement

// Launch-error to status code; never throws across the C ABI.
#define DF_RETURN_LAST_ERROR()                                   \
    do {                                                         \
        cudaError_t e__ = cudaGetLastError();                    \
        return e__ == cudaSuccess ? DF_OK : (int)e__;            \
    } while (0)

namespace df {

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Squared distance with the reference kernel's rounding order
// (lib/knn/src/knn_cuda_kernel.cu:80-84 compiled with default -fmad=true):
//   ssd = fma(dz,dz, fma(dy,dy, fma(dx,dx,0)))   and fma(dx,dx,0) == rn(dx*dx).
// Intrinsics pin the rounding so the optimiser can neither fuse nor split anything.
__device__ __forceinline__ float ref_ssd3(float rx, float ry, float rz, float qx, float qy, float qz)
{
    const float dx = __fsub_rn(rx, qx);
    const float dy = __fsub_rn(ry, qy);
    const float dz = __fsub_rn(rz, qz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// Rotation matrix of a UNIT quaternion (w,x,y,z), term order of lib/loss.py:18-26.
__device__ __forceinline__ void quat_to_rot(float w, float x, float y, float z, float (&R)[9])
{
    R[0] = 1.0f - 2.0f * (y * y + z * z);
    R[1] = 2.0f * x * y - 2.0f * w * z;
    R[2] = 2.0f * w * y + 2.0f * x * z;
    R[3] = 2.0f * x * y + 2.0f * z * w;
    R[4] = 1.0f - 2.0f * (x * x + z * z);
    R[5] = -2.0f * w * x + 2.0f * y * z;
    R[6] = -2.0f * w * y + 2.0f * x * z;
    R[7] = 2.0f * w * x + 2.0f * y * z;
    R[8] = 1.0f - 2.0f * (x * x + y * y);
}

}  // namespace df
