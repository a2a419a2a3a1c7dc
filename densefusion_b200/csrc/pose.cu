// K5 -- on-device pose selection, cloud re-expression and pose composition for the iterative
// refinement loop (reference: tools/eval_ycb.py:193-233, with lib/transformations.py:1254-1278
// quaternion_matrix and :1281-1363 quaternion_from_matrix(isprecise=True)).
//
// The reference leaves the GPU 2 + 2*iters times per object (argmax -> numpy float64 4x4 algebra ->
// back).  Here the pose state is a float64[7] per crop that never leaves HBM: the float64 algebra is
// done by one thread per crop, the O(N) cloud transform by one CTA per crop, and the whole
// estimate -> select -> (transform -> refine -> compose) x iters chain is graph-capturable.
#include "df_common.cuh"
#include "../../include/densefusion_b200.h"
#include <math_constants.h>

namespace {

__device__ void quaternion_matrix_f64(const double* q_in, double (&R)[9])
{
    double q[4] = {q_in[0], q_in[1], q_in[2], q_in[3]};
    const double n = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    if (n < 2.220446049250313e-16 * 4.0) {               // _EPS, transformations.py:1270
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        return;
    }
    const double s = sqrt(2.0 / n);
    for (int i = 0; i < 4; ++i) q[i] *= s;
    R[0] = 1.0 - q[2] * q[2] - q[3] * q[3]; R[1] = q[1] * q[2] - q[3] * q[0]; R[2] = q[1] * q[3] + q[2] * q[0];
    R[3] = q[1] * q[2] + q[3] * q[0]; R[4] = 1.0 - q[1] * q[1] - q[3] * q[3]; R[5] = q[2] * q[3] - q[1] * q[0];
    R[6] = q[1] * q[3] - q[2] * q[0]; R[7] = q[2] * q[3] + q[1] * q[0]; R[8] = 1.0 - q[1] * q[1] - q[2] * q[2];
}

// isprecise=True branch; M[3][3] == 1 for a homogeneous rotation
__device__ void quaternion_from_matrix_f64(const double (&M)[9], double* q)
{
    double t = M[0] + M[4] + M[8] + 1.0;                 // numpy.trace of the 4x4
    if (t > 1.0) {
        q[0] = t;
        q[3] = M[3] - M[1];
        q[2] = M[2] - M[6];
        q[1] = M[7] - M[5];
    } else {
        int i = 0, j = 1, k = 2;
        if (M[4] > M[0]) { i = 1; j = 2; k = 0; }
        if (M[8] > M[i * 3 + i]) { i = 2; j = 0; k = 1; }
        t = M[i * 3 + i] - (M[j * 3 + j] + M[k * 3 + k]) + 1.0;
        double v[4];
        v[i] = t;
        v[j] = M[i * 3 + j] + M[j * 3 + i];
        v[k] = M[k * 3 + i] + M[i * 3 + k];
        v[3] = M[k * 3 + j] - M[j * 3 + k];
        q[0] = v[3]; q[1] = v[0]; q[2] = v[1]; q[3] = v[2];
    }
    const double s = 0.5 / sqrt(t * 1.0);
    for (int c = 0; c < 4; ++c) q[c] *= s;
    if (q[0] < 0.0)
        for (int c = 0; c < 4; ++c) q[c] = -q[c];
}

// first-index argmax of the confidences, q/|q| in fp32, t = point + offset  (eval_ycb.py:193-201)
__global__ void __launch_bounds__(128)
select_pose_kernel(const float* __restrict__ pred_r, const float* __restrict__ pred_t, const float* __restrict__ pred_c,
                   const float* __restrict__ points, int N, double* __restrict__ pose, int64_t* __restrict__ which)
{
    __shared__ float s_c[4];
    __shared__ int s_k[4];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* conf = pred_c + (size_t)b * N;
    float cb = -CUDART_INF_F;
    int kb = 0x7fffffff;
    for (int h = tid; h < N; h += 128) {
        const float c = conf[h];
        if (c > cb) { cb = c; kb = h; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float oc = __shfl_xor_sync(0xffffffffu, cb, o);
        const int ok = __shfl_xor_sync(0xffffffffu, kb, o);
        if (oc > cb || (oc == cb && ok < kb)) { cb = oc; kb = ok; }
    }
    if ((tid & 31) == 0) { s_c[tid >> 5] = cb; s_k[tid >> 5] = kb; }
    __syncthreads();
    if (tid == 0) {
        for (int wv = 1; wv < 4; ++wv)
            if (s_c[wv] > cb || (s_c[wv] == cb && s_k[wv] < kb)) { cb = s_c[wv]; kb = s_k[wv]; }
        if (kb == 0x7fffffff) kb = 0;
        const float* q = pred_r + ((size_t)b * N + kb) * 4;
        const float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        double* o = pose + (size_t)b * 7;
        for (int c = 0; c < 4; ++c) o[c] = (double)(q[c] / n);
        for (int c = 0; c < 3; ++c)
            o[4 + c] = (double)(points[((size_t)b * N + kb) * 3 + c] + pred_t[((size_t)b * N + kb) * 3 + c]);
        if (which) which[b] = kb;
    }
}

// new_cloud = (cloud - T) . R with R, T rounded to fp32 like `.astype(np.float32)` (eval_ycb.py:206-211)
__global__ void __launch_bounds__(256)
cloud_transform_kernel(const float* __restrict__ cloud, const double* __restrict__ pose, float* __restrict__ out, int N)
{
    __shared__ float s_R[9], s_T[3];
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        double R[9];
        quaternion_matrix_f64(pose + (size_t)b * 7, R);
        for (int i = 0; i < 9; ++i) s_R[i] = (float)R[i];
        for (int i = 0; i < 3; ++i) s_T[i] = (float)pose[(size_t)b * 7 + 4 + i];
    }
    __syncthreads();
    const float* src = cloud + (size_t)b * N * 3;
    float* dst = out + (size_t)b * N * 3;
    for (int n = threadIdx.x; n < N; n += 256) {
        const float vx = src[n * 3] - s_T[0], vy = src[n * 3 + 1] - s_T[1], vz = src[n * 3 + 2] - s_T[2];
        dst[n * 3 + 0] = vx * s_R[0] + vy * s_R[3] + vz * s_R[6];
        dst[n * 3 + 1] = vx * s_R[1] + vy * s_R[4] + vz * s_R[7];
        dst[n * 3 + 2] = vx * s_R[2] + vy * s_R[5] + vz * s_R[8];
    }
}

// pose <- pose o (normalised r2, t2)   (eval_ycb.py:213-229), float64
__global__ void pose_compose_kernel(double* __restrict__ pose, const float* __restrict__ r2, const float* __restrict__ t2, int B)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double* ps = pose + (size_t)b * 7;
    const float* q2f = r2 + (size_t)b * 4;
    const float n = sqrtf(q2f[0] * q2f[0] + q2f[1] * q2f[1] + q2f[2] * q2f[2] + q2f[3] * q2f[3]);
    double q2[4];
    for (int c = 0; c < 4; ++c) q2[c] = (double)(q2f[c] / n);
    double R1[9], R2[9], F[9];
    quaternion_matrix_f64(ps, R1);
    quaternion_matrix_f64(q2, R2);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            F[i * 3 + j] = R1[i * 3] * R2[j] + R1[i * 3 + 1] * R2[3 + j] + R1[i * 3 + 2] * R2[6 + j];
    double tn[3];
    for (int i = 0; i < 3; ++i)
        tn[i] = R1[i * 3] * (double)t2[b * 3] + R1[i * 3 + 1] * (double)t2[b * 3 + 1] + R1[i * 3 + 2] * (double)t2[b * 3 + 2] + ps[4 + i];
    double qn[4];
    quaternion_from_matrix_f64(F, qn);
    for (int c = 0; c < 4; ++c) ps[c] = qn[c];
    for (int c = 0; c < 3; ++c) ps[4 + c] = tn[c];
}

}  // namespace

extern "C" int df_select_pose(const float* pred_r, const float* pred_t, const float* pred_c, const float* points,
                              int B, int N, double* pose, int64_t* which, void* stream)
{
    if (!pred_r || !pred_t || !pred_c || !points || !pose || B <= 0 || N <= 0) return DF_ERR_ARG;
    select_pose_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(pred_r, pred_t, pred_c, points, N, pose, which);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_cloud_transform(const float* cloud, const double* pose, float* out, int B, int N, void* stream)
{
    if (!cloud || !pose || !out || B <= 0 || N <= 0) return DF_ERR_ARG;
    cloud_transform_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(cloud, pose, out, N);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_pose_compose(double* pose, const float* r2, const float* t2, int B, void* stream)
{
    if (!pose || !r2 || !t2 || B <= 0) return DF_ERR_ARG;
    pose_compose_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(pose, r2, t2, B);
    DF_RETURN_LAST_ERROR();
}
