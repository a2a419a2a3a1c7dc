// Shared helpers for the densefusion_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math_constants.h>

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

// ---- bit-exact 1-NN scan over a shared-memory tile of reference points -----------------------------
// The reference keeps a running (min, index) with one compare + two selects per pair.  Here the refs are
// walked in chunks of NN_CHUNK: per pair only the distance (6 fp32 ops) and a share of an FMNMX tree is paid,
// and the running state is (min, first chunk that attained it).  The index inside the winning chunk is
// recovered afterwards by nn_resolve().  Semantics are unchanged: strict '<' between chunks keeps the EARLIEST
// chunk, nn_resolve takes the FIRST equal element, fminf drops NaN candidates exactly like `NaN < best` being
// false, and a NaN seed (row 0) never loses, as in knn_cuda_kernel.cu:122,150-167.
constexpr int NN_CHUNK = 8;

__device__ __forceinline__ float fmin3(float a, float b, float c)
{
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// s_ref must be padded to a multiple of NN_CHUNK with (+inf,+inf,+inf) entries.
template <int QPT>
__device__ __forceinline__ void nn_scan_tile(const float4* __restrict__ s_ref, int n_padded, int base,
                                             const float (&qx)[QPT], const float (&qy)[QPT], const float (&qz)[QPT],
                                             float (&best)[QPT], int (&best_chunk)[QPT])
{
#pragma unroll 1
    for (int r = 0; r < n_padded; r += NN_CHUNK) {
        float d[QPT][NN_CHUNK];
#pragma unroll
        for (int c = 0; c < NN_CHUNK; ++c) {
            const float4 p = s_ref[r + c];
#pragma unroll
            for (int i = 0; i < QPT; ++i) d[i][c] = ref_ssd3(p.x, p.y, p.z, qx[i], qy[i], qz[i]);
        }
#pragma unroll
        for (int i = 0; i < QPT; ++i) {
            // FMNMX3 tree (sm_100): 4 instructions for 8 candidates; NaN candidates are dropped
            float m = fmin3(d[i][0], d[i][1], d[i][2]);
            m = fmin3(m, d[i][3], d[i][4]);
            m = fmin3(m, d[i][5], d[i][6]);
            m = fminf(m, d[i][7]);
            if (m < best[i]) { best[i] = m; best_chunk[i] = base + r; }
        }
    }
}

// First index in [chunk, chunk + NN_CHUNK) (clipped to R) whose distance equals `best`; 0 if none (NaN seed).
// ref points are read from global memory in dim-major (x[], y[], z[]) or point-major (xyz xyz ...) layout.
template <bool POINT_MAJOR>
__device__ __forceinline__ int nn_resolve(const float* __restrict__ ref, int R, int chunk, float best,
                                          float qx, float qy, float qz)
{
    int found = -1;
#pragma unroll
    for (int c = NN_CHUNK - 1; c >= 0; --c) {
        const int r = chunk + c;
        if (r < R) {
            const float rx = POINT_MAJOR ? ref[r * 3] : ref[r];
            const float ry = POINT_MAJOR ? ref[r * 3 + 1] : ref[(size_t)R + r];
            const float rz = POINT_MAJOR ? ref[r * 3 + 2] : ref[2 * (size_t)R + r];
            if (ref_ssd3(rx, ry, rz, qx, qy, qz) == best) found = r;
        }
    }
    return found < 0 ? 0 : found;
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
