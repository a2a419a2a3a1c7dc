// Backward-pass kernels of the dense-fusion head / refiner (what autograd derives from lib/network.py) and the
// optimiser step of the data-parallel training loop (tools/train.py:152-169).  Exact fp32.
//   df_gemm_wgrad_fp32        dW[g][n,k] = sum_m dY[m, g*dy_gs + n] * X[m, g*x_gs + k]   (split over m, deterministic)
//   df_reduce_partials        out (+)= sum_s part[s]                                       (fixed order)
//   df_colsum_rows            out[group, c] = sum_{rows of group} X[row, c]                (bias grads, per-crop sums)
//   df_relu_mask_inplace      d *= (act > 0)
//   df_pool_backward          dh[row,c] = dg[crop,c] / N * (h[row,c] > 0)                  (AvgPool1d + ReLU backward)
//   df_select_out_backward    last tower layer for the selected object: dh, and per-crop weight / bias grads
//   df_accumulate_selected    scatter the per-crop (8 x 128) blocks into the (num_obj*w, 128) weight grads in crop order
//   df_gather_embedding_backward   dfeat[b,c,choose[b,n]] += demb[b*N+n, c]
//   df_adam_step              torch.optim.Adam's update on a flat arena
#include "df_common.cuh"
#include "../../include/densefusion_b200.h"

namespace {

constexpr int WG_BM = 16;     // rows (reduction) per smem stage

// 128 x 128 output tile per CTA, 8x8 per thread, reduction over a slice of the rows
__global__ void __launch_bounds__(256)
wgrad_kernel(const float* __restrict__ dY, int ldy, const float* __restrict__ X, int ldx, float* __restrict__ part,
             int M, int N, int K, int splits, long long dy_gs, long long x_gs)
{
    __shared__ __align__(16) float Ys[WG_BM][128 + 4];
    __shared__ __align__(16) float Xs[WG_BM][128 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int g = blockIdx.z / splits, sp = blockIdx.z - g * splits;
    const int n0 = blockIdx.x * 128, k0 = blockIdx.y * 128;
    dY += g * dy_gs;
    X += g * x_gs;
    const int rows_per_split = ((M + splits - 1) / splits + WG_BM - 1) / WG_BM * WG_BM;
    const int m_begin = sp * rows_per_split, m_end = min(M, m_begin + rows_per_split);
    const bool vec_y = (ldy % 4 == 0) && (((uintptr_t)dY & 15) == 0) && (n0 + 128 <= N);
    const bool vec_x = (ldx % 4 == 0) && (((uintptr_t)X & 15) == 0) && (k0 + 128 <= K);

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    for (int m0 = m_begin; m0 < m_end; m0 += WG_BM) {
        // 16 x 128 floats each: 512 float4, 2 per thread
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int f = tid + i * 256;
            const int r = f >> 5, c4 = (f & 31) * 4;
            const int m = m0 + r;
            float4 vy = make_float4(0.f, 0.f, 0.f, 0.f), vx = vy;
            if (m < m_end) {
                if (vec_y) vy = __ldg(reinterpret_cast<const float4*>(dY + (size_t)m * ldy + n0 + c4));
                else {
                    const float* q = dY + (size_t)m * ldy;
                    vy.x = n0 + c4 + 0 < N ? q[n0 + c4 + 0] : 0.f; vy.y = n0 + c4 + 1 < N ? q[n0 + c4 + 1] : 0.f;
                    vy.z = n0 + c4 + 2 < N ? q[n0 + c4 + 2] : 0.f; vy.w = n0 + c4 + 3 < N ? q[n0 + c4 + 3] : 0.f;
                }
                if (vec_x) vx = __ldg(reinterpret_cast<const float4*>(X + (size_t)m * ldx + k0 + c4));
                else {
                    const float* q = X + (size_t)m * ldx;
                    vx.x = k0 + c4 + 0 < K ? q[k0 + c4 + 0] : 0.f; vx.y = k0 + c4 + 1 < K ? q[k0 + c4 + 1] : 0.f;
                    vx.z = k0 + c4 + 2 < K ? q[k0 + c4 + 2] : 0.f; vx.w = k0 + c4 + 3 < K ? q[k0 + c4 + 3] : 0.f;
                }
            }
            *reinterpret_cast<float4*>(&Ys[r][c4]) = vy;
            *reinterpret_cast<float4*>(&Xs[r][c4]) = vx;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < WG_BM; ++r) {
            float a[8], b[8];
            const float4 a0 = *reinterpret_cast<const float4*>(&Ys[r][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&Ys[r][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Xs[r][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Xs[r][64 + tx * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* out = part + ((size_t)sp * gridDim.z / splits + g) * (size_t)N * K;    // [split][group][N][K]
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = n0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (k < K) out[(size_t)n * K + k] = acc[i][j];
        }
    }
}

__global__ void reduce_partials_kernel(const float* __restrict__ part, int splits, long long count,
                                       float* __restrict__ out, int accumulate)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float s = 0.0f;
    for (int sp = 0; sp < splits; ++sp) s += part[(size_t)sp * count + i];
    out[i] = accumulate ? out[i] + s : s;
}

// Column sums over the rows of a group.  One CTA per (32-column slab, group): 8 float4 column lanes x 128 row lanes, so
// a (8000 x 1920) bias-gradient sum runs on 60 CTAs x 1024 threads with 16-byte loads instead of 15 CTAs walking
// 4000 rows each (that first version was 12-31% of a training step).  float64 accumulation: these sums add thousands of
// terms of mixed sign and the kernel is bandwidth-bound, so the wider adds are free.  Fixed-order tree -> deterministic.
__global__ void __launch_bounds__(1024)
colsum_rows_kernel(const float* __restrict__ X, int ldx, int rows_per_group, int C, float* __restrict__ out, int accumulate)
{
    __shared__ double red[128][33];
    const int grp = blockIdx.y, cq = threadIdx.x & 7, rl = threadIdx.x >> 3;
    const int c = blockIdx.x * 32 + cq * 4;
    const float* base = X + (size_t)grp * rows_per_group * ldx;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (c < C) {
        if (c + 3 < C && (ldx & 3) == 0 && (((uintptr_t)X) & 15) == 0) {
#pragma unroll 4
            for (int r = rl; r < rows_per_group; r += 128) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)r * ldx + c));
                s0 += (double)v.x; s1 += (double)v.y; s2 += (double)v.z; s3 += (double)v.w;
            }
        } else {
            for (int r = rl; r < rows_per_group; r += 128) {
                const float* q = base + (size_t)r * ldx + c;
                s0 += (double)q[0];
                if (c + 1 < C) s1 += (double)q[1];
                if (c + 2 < C) s2 += (double)q[2];
                if (c + 3 < C) s3 += (double)q[3];
            }
        }
    }
    red[rl][cq * 4 + 0] = s0; red[rl][cq * 4 + 1] = s1; red[rl][cq * 4 + 2] = s2; red[rl][cq * 4 + 3] = s3;
    __syncthreads();
    for (int half = 64; half >= 1; half >>= 1) {
        if (rl < half) {
#pragma unroll
            for (int k = 0; k < 4; ++k) red[rl][cq * 4 + k] += red[rl + half][cq * 4 + k];
        }
        __syncthreads();
    }
    if (threadIdx.x < 32) {
        const int cc = blockIdx.x * 32 + threadIdx.x;
        if (cc < C) {
            const float t = (float)red[0][threadIdx.x];
            float* o = out + (size_t)grp * C + cc;
            *o = accumulate ? *o + t : t;
        }
    }
}

__global__ void relu_mask_kernel(float* __restrict__ d, const float* __restrict__ act, int ld, int cols, long long rows)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = rows * cols;
    if (i >= total) return;
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    const size_t o = (size_t)r * ld + c;
    if (!(act[o] > 0.0f)) d[o] = 0.0f;
}

__global__ void pool_backward_kernel(const float* __restrict__ dg, const float* __restrict__ h, float* __restrict__ dh,
                                     int rows_per_crop, int C, long long rows)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * C) return;
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const float v = dg[(r / rows_per_crop) * C + c] / (float)rows_per_crop;
    dh[i] = h[i] > 0.0f ? v : 0.0f;
}

// one warp per row: dh[row, branch*128 + k] = relu'(h) * sum_i g[row,i] W[obj*w+i, k]; c branch goes through sigmoid'
__global__ void __launch_bounds__(256)
select_out_dh_kernel(const float* __restrict__ g_r, const float* __restrict__ g_t, const float* __restrict__ g_c,
                     const float* __restrict__ out_c, const float* __restrict__ h, int ldh,
                     const float* __restrict__ Wr, const float* __restrict__ Wt, const float* __restrict__ Wc,
                     const int64_t* __restrict__ obj, int rows_per_crop, int num_obj, long long rows,
                     float* __restrict__ dh, float* __restrict__ gz /* (rows,8) pre-activation grads */)
{
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    long long o = obj[row / rows_per_crop];
    o = o < 0 ? 0 : (o >= num_obj ? num_obj - 1 : o);
    float g[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) g[i] = g_r ? g_r[row * 4 + i] : 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) g[4 + i] = g_t ? g_t[row * 3 + i] : 0.0f;
    g[7] = 0.0f;
    if (Wc && g_c) { const float c = out_c[row]; g[7] = g_c[row] * c * (1.0f - c); }
    if (lane < 8) gz[row * 8 + lane] = g[lane];
    const float* hr = h + row * ldh;
    float* dr = dh + row * ldh;
    const int nb = Wc ? 3 : 2;
    for (int br = 0; br < nb; ++br) {
        const float4 hv = *reinterpret_cast<const float4*>(hr + br * 128 + lane * 4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int w = br == 0 ? 4 : (br == 1 ? 3 : 1);
        const int g0 = br == 0 ? 0 : (br == 1 ? 4 : 7);
        const float* W = br == 0 ? Wr : (br == 1 ? Wt : Wc);
        for (int i = 0; i < w; ++i) {
            const float4 wv = __ldg(reinterpret_cast<const float4*>(W + (o * w + i) * 128 + lane * 4));
            acc.x += g[g0 + i] * wv.x; acc.y += g[g0 + i] * wv.y; acc.z += g[g0 + i] * wv.z; acc.w += g[g0 + i] * wv.w;
        }
        acc.x = hv.x > 0.f ? acc.x : 0.f; acc.y = hv.y > 0.f ? acc.y : 0.f;
        acc.z = hv.z > 0.f ? acc.z : 0.f; acc.w = hv.w > 0.f ? acc.w : 0.f;
        *reinterpret_cast<float4*>(dr + br * 128 + lane * 4) = acc;
    }
}

// one CTA per (crop, row slice), thread k: blk[crop][slice][i][k] = sum over the slice's rows of gz[row,i] * h[row, branch(i)*128 + k];
// bsum[crop][slice][i] = sum gz.  DF_SELECT_SPLITS slices per crop so that 16 crops still fill the machine.
__global__ void __launch_bounds__(128)
select_out_wgrad_kernel(const float* __restrict__ gz, const float* __restrict__ h, int ldh, int rows_per_crop,
                        float* __restrict__ blk, float* __restrict__ bsum)
{
    const int crop = blockIdx.x, sp = blockIdx.y, k = threadIdx.x;
    const int per = (rows_per_crop + DF_SELECT_SPLITS - 1) / DF_SELECT_SPLITS;
    const int r0 = sp * per, r1 = min(rows_per_crop, r0 + per);
    float acc[8], bs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] = 0.f; bs[i] = 0.f; }
    for (int r = r0; r < r1; ++r) {
        const long long row = (long long)crop * rows_per_crop + r;
        const float hr = h[row * ldh + k], ht = h[row * ldh + 128 + k];
        const float hc = ldh >= 384 ? h[row * ldh + 256 + k] : 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float g = gz[row * 8 + i];
            acc[i] += g * (i < 4 ? hr : (i < 7 ? ht : hc));
            bs[i] += g;
        }
    }
    const size_t slot = (size_t)crop * DF_SELECT_SPLITS + sp;
#pragma unroll
    for (int i = 0; i < 8; ++i) blk[(slot * 8 + i) * 128 + k] = acc[i];
    if (k < 8) bsum[slot * 8 + k] = bs[k];
}

// single CTA, crops and slices in ascending order (deterministic even when several crops share an object)
__global__ void __launch_bounds__(128)
accumulate_selected_kernel(const float* __restrict__ blk, const float* __restrict__ bsum, const int64_t* __restrict__ obj,
                           int crops, int num_obj, float* dWr, float* dbr, float* dWt, float* dbt, float* dWc, float* dbc)
{
    const int k = threadIdx.x;
    for (int c = 0; c < crops; ++c) {
        long long o = obj[c];
        o = o < 0 ? 0 : (o >= num_obj ? num_obj - 1 : o);
        float v[8], b = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        for (int sp = 0; sp < DF_SELECT_SPLITS; ++sp) {
            const size_t slot = (size_t)c * DF_SELECT_SPLITS + sp;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += blk[(slot * 8 + i) * 128 + k];
            if (k < 8) b += bsum[slot * 8 + k];
        }
        for (int i = 0; i < 8; ++i) {
            if (i < 4) dWr[(o * 4 + i) * 128 + k] += v[i];
            else if (i < 7) dWt[(o * 3 + (i - 4)) * 128 + k] += v[i];
            else if (dWc) dWc[o * 128 + k] += v[i];
        }
        if (k < 8) {
            if (k < 4) dbr[o * 4 + k] += b;
            else if (k < 7) dbt[o * 3 + (k - 4)] += b;
            else if (dbc) dbc[o] += b;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
gather_backward_kernel(const float* __restrict__ demb, const int64_t* __restrict__ choose, float* __restrict__ dfeat,
                       long long sb, long long sc, long long sp, int B, int N, int HW)
{
    const int lane = threadIdx.x & 31;
    const long long pt = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pt >= (long long)B * N) return;
    const int b = (int)(pt / N);
    long long pix = choose[pt];
    pix = pix < 0 ? 0 : (pix >= HW ? HW - 1 : pix);
    atomicAdd(dfeat + b * sb + lane * sc + pix * sp, demb[pt * 32 + lane]);
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.0f - b1) * gi;          // exp_avg.lerp_(grad, 1-beta1)
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;     // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
}

// CUDA-graph friendly form: the step number lives in device memory (a captured launch cannot change its scalar arguments)
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                long long n, float lr, float b1, float b2, float eps, const int* __restrict__ step_ptr)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // bias corrections in double like torch.optim.Adam (python floats): fp32 powf deviates by ~1e-5 relative in the first steps
    const double step = (double)(*step_ptr + 1);
    const float bc1 = (float)(1.0 - pow((double)b1, step));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, step));
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
}

__global__ void bump_kernel(int* step_ptr) { *step_ptr += 1; }

}  // namespace

extern "C" int df_gemm_wgrad_fp32(const float* dY, int ldy, const float* X, int ldx, float* partial, int M, int N, int K,
                                  int groups, int splits, long long dy_group_stride, long long x_group_stride, void* stream)
{
    if (!dY || !X || !partial || M <= 0 || N <= 0 || K <= 0 || groups <= 0 || splits <= 0) return DF_ERR_ARG;
    if ((long long)groups * splits > 65535) return DF_ERR_ARG;
    dim3 grid((N + 127) / 128, (K + 127) / 128, groups * splits);
    wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dY, ldy, X, ldx, partial, M, N, K, splits, dy_group_stride,
                                                         x_group_stride);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_reduce_partials(const float* partial, int splits, long long count, float* out, int accumulate, void* stream)
{
    if (!partial || !out || splits <= 0 || count <= 0) return DF_ERR_ARG;
    reduce_partials_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(partial, splits, count, out, accumulate);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_colsum_rows(const float* X, int ldx, int rows_per_group, int groups, int C, float* out, int accumulate,
                              void* stream)
{
    if (!X || !out || rows_per_group <= 0 || groups <= 0 || groups > 65535 || C <= 0) return DF_ERR_ARG;
    colsum_rows_kernel<<<dim3((C + 31) / 32, groups), 1024, 0, (cudaStream_t)stream>>>(X, ldx, rows_per_group, C, out, accumulate);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_relu_mask_inplace(float* d, const float* act, int ld, int cols, long long rows, void* stream)
{
    if (!d || !act || ld <= 0 || cols <= 0 || rows <= 0) return DF_ERR_ARG;
    const long long total = rows * cols;
    relu_mask_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d, act, ld, cols, rows);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_pool_backward(const float* dg, const float* h, float* dh, int rows_per_crop, int C, long long rows,
                                void* stream)
{
    if (!dg || !h || !dh || rows_per_crop <= 0 || C <= 0 || rows <= 0) return DF_ERR_ARG;
    const long long total = rows * C;
    pool_backward_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dg, h, dh, rows_per_crop, C, rows);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_select_out_backward(const float* g_r, const float* g_t, const float* g_c, const float* out_c,
                                      const float* h, int ldh, const float* Wr, const float* Wt, const float* Wc,
                                      const int64_t* obj, int rows_per_crop, int num_obj, long long rows, float* dh,
                                      float* gz, float* blk, float* bsum, float* dWr, float* dbr, float* dWt, float* dbt,
                                      float* dWc, float* dbc, void* stream)
{
    if (!h || !Wr || !Wt || !obj || !dh || !gz || !blk || !bsum || !dWr || !dbr || !dWt || !dbt) return DF_ERR_ARG;
    if (rows <= 0 || rows_per_crop <= 0 || rows % rows_per_crop || num_obj <= 0 || ldh % 4 || ldh < (Wc ? 384 : 256)) return DF_ERR_ARG;
    if (Wc && (!out_c || !dWc || !dbc)) return DF_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int crops = (int)(rows / rows_per_crop);
    select_out_dh_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(g_r, g_t, g_c, out_c, h, ldh, Wr, Wt, Wc, obj,
                                                                    rows_per_crop, num_obj, rows, dh, gz);
    select_out_wgrad_kernel<<<dim3(crops, DF_SELECT_SPLITS), 128, 0, s>>>(gz, h, ldh, rows_per_crop, blk, bsum);
    accumulate_selected_kernel<<<1, 128, 0, s>>>(blk, bsum, obj, crops, num_obj, dWr, dbr, dWt, dbt, dWc, dbc);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_gather_embedding_backward(const float* demb, const int64_t* choose, float* dfeat, long long stride_b,
                                            long long stride_c, long long stride_pix, int B, int N, int HW, void* stream)
{
    if (!demb || !choose || !dfeat || B <= 0 || N <= 0 || HW <= 0) return DF_ERR_ARG;
    const long long pts = (long long)B * N;
    gather_backward_kernel<<<(unsigned)((pts + 7) / 8), 256, 0, (cudaStream_t)stream>>>(demb, choose, dfeat, stride_b,
                                                                                       stride_c, stride_pix, B, N, HW);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                            float beta1, float beta2, float eps, int step, void* stream)
{
    if (!param || !grad || !exp_avg || !exp_avg_sq || n <= 0 || step <= 0) return DF_ERR_ARG;
    const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1,
                                                                             beta2, eps, bc1, bc2_sqrt);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                                float beta1, float beta2, float eps, int* step_counter, void* stream)
{
    if (!param || !grad || !exp_avg || !exp_avg_sq || !step_counter || n <= 0) return DF_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    adam_dev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                               step_counter);
    bump_kernel<<<1, 1, 0, s>>>(step_counter);
    DF_RETURN_LAST_ERROR();
}
