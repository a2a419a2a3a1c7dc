// fp32 SIMT GEMM with fused epilogues -- the exact-fp32 arithmetic mode of the dense-fusion head
// (reference: the Conv1d(k=1) / Linear layers of lib/network.py:54-63, :107-121, :151-166, :193-200).
//
//   C[m, n] = act( sum_k A[m, k] * W[n, k] + bias[n] )          A: points x channels (K contiguous)
//                                                               W: torch (out, in) weight, K contiguous
// Epilogue variants (all fused, no extra pass over C):
//   * per-crop bias: bias row selected by m / rows_per_crop -- carries the folded global feature
//     (W1[:, 384:1408] . avgpool) into conv1_{r,t,c} instead of broadcasting a 1024-wide tensor.
//   * column-sum pooling: the (points x 1024) conv6 activation is never stored; each CTA writes the
//     column sums of its crop-aligned row tile and df_pool_finish adds the tiles in a fixed order.
//   * groups on blockIdx.z: the r / t / c towers as one launch (block-diagonal weights).
// FFMA-pipe bound: 128x128x16 tiles, 8x8 outputs per thread, float4 LDS, register-prefetch double
// buffering.  This is the parity mode; gemm_tc.cu is the tcgen05 3xTF32 mode of the same contract.
#include "df_common.cuh"
#include "../../include/densefusion_b200.h"

namespace {

struct GemmParams {
    const float* A; int lda;
    const float* W; int ldw;
    const float* bias; int bias_crop_stride;
    float* C; int ldc;
    int M, N, K, relu;
    int rows_per_crop;
    long long a_gs, w_gs, bias_gs, c_gs;
    float* pool_partial; int tiles_per_crop;
    const float* relu_mask;       // backward: multiply by (mask[m,n] > 0); same ld / group stride as C
    int accumulate;               // backward: C += result
};

constexpr int BK = 16;

template <int BM, int BN, int TMG, int TNG>
__global__ void __launch_bounds__((BM / (4 * TMG)) * (BN / (4 * TNG)))
sgemm_kernel(const GemmParams p)
{
    constexpr int NX = BN / (4 * TNG), NY = BM / (4 * TMG), NT = NX * NY;
    constexpr int A_F4 = BM * BK / 4, B_F4 = BN * BK / 4;
    constexpr int A_PER = (A_F4 + NT - 1) / NT, B_PER = (B_F4 + NT - 1) / NT;
    constexpr int TM = 4 * TMG, TN = 4 * TNG;
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int tid = threadIdx.x, tx = tid % NX, ty = tid / NX;
    const int g = blockIdx.z;
    const float* A = p.A + g * p.a_gs;
    const float* W = p.W + g * p.w_gs;
    float* C = p.C ? p.C + g * p.c_gs : nullptr;
    const int n0 = blockIdx.x * BN;

    // row tile: plain, or aligned to a crop when pooling
    int row0, rows_valid, crop = 0, tile_in_crop = 0;
    if (p.pool_partial) {
        crop = blockIdx.y / p.tiles_per_crop;
        tile_in_crop = blockIdx.y - crop * p.tiles_per_crop;
        row0 = crop * p.rows_per_crop + tile_in_crop * BM;
        rows_valid = min(BM, p.rows_per_crop - tile_in_crop * BM);
    } else {
        row0 = blockIdx.y * BM;
        rows_valid = min(BM, p.M - row0);
    }

    float4 ra[A_PER], rb[B_PER];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int f = tid + i * NT;
            const int r = f >> 2, kq = f & 3;
            ra[i] = (f < A_F4 && r < rows_valid)
                        ? __ldg(reinterpret_cast<const float4*>(A + (size_t)(row0 + r) * p.lda + k0 + kq * 4))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            const int f = tid + i * NT;
            const int r = f >> 2, kq = f & 3;
            rb[i] = (f < B_F4 && n0 + r < p.N)
                        ? __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + r) * p.ldw + k0 + kq * 4))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int f = tid + i * NT;
            if (f < A_F4) {
                const int r = f >> 2, kq = (f & 3) * 4;
                As[buf][kq + 0][r] = ra[i].x; As[buf][kq + 1][r] = ra[i].y;
                As[buf][kq + 2][r] = ra[i].z; As[buf][kq + 3][r] = ra[i].w;
            }
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            const int f = tid + i * NT;
            if (f < B_F4) {
                const int r = f >> 2, kq = (f & 3) * 4;
                Bs[buf][kq + 0][r] = rb[i].x; Bs[buf][kq + 1][r] = rb[i].y;
                Bs[buf][kq + 2][r] = rb[i].z; Bs[buf][kq + 3][r] = rb[i].w;
            }
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    fetch(0);
    stash(0);
    __syncthreads();
    const int nk = p.K / BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) fetch((kt + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int gi = 0; gi < TMG; ++gi) {
                const float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][gi * (BM / TMG) + ty * 4]);
                a[gi * 4 + 0] = v.x; a[gi * 4 + 1] = v.y; a[gi * 4 + 2] = v.z; a[gi * 4 + 3] = v.w;
            }
#pragma unroll
            for (int gj = 0; gj < TNG; ++gj) {
                const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][gj * (BN / TNG) + tx * 4]);
                b[gj * 4 + 0] = v.x; b[gj * 4 + 1] = v.y; b[gj * 4 + 2] = v.z; b[gj * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            stash(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- epilogue: bias, activation, store or pool ----
    const float* bias = p.bias ? p.bias + g * p.bias_gs : nullptr;
    float colsum[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) colsum[j] = 0.0f;
#pragma unroll
    for (int gi = 0; gi < TMG; ++gi) {
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int i = gi * 4 + ii;
            const int r = gi * (BM / TMG) + ty * 4 + ii;
            if (r >= rows_valid) continue;
            const int row = row0 + r;
            const float* brow = bias;
            if (bias && p.bias_crop_stride) brow = bias + (size_t)(row / p.rows_per_crop) * p.bias_crop_stride;
#pragma unroll
            for (int gj = 0; gj < TNG; ++gj) {
                const int col = n0 + gj * (BN / TNG) + tx * 4;
                if (col >= p.N) continue;
                float4 v = make_float4(acc[i][gj * 4], acc[i][gj * 4 + 1], acc[i][gj * 4 + 2], acc[i][gj * 4 + 3]);
                if (brow) {
                    const float4 bv = __ldg(reinterpret_cast<const float4*>(brow + col));
                    v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                }
                if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                if (p.relu_mask) {
                    const float4 mk = __ldg(reinterpret_cast<const float4*>(p.relu_mask + g * p.c_gs + (size_t)row * p.ldc + col));
                    v.x = mk.x > 0.f ? v.x : 0.f; v.y = mk.y > 0.f ? v.y : 0.f;
                    v.z = mk.z > 0.f ? v.z : 0.f; v.w = mk.w > 0.f ? v.w : 0.f;
                }
                if (p.accumulate) {
                    const float4 old = *reinterpret_cast<const float4*>(C + (size_t)row * p.ldc + col);
                    v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
                }
                if (p.pool_partial) {
                    colsum[gj * 4] += v.x; colsum[gj * 4 + 1] += v.y; colsum[gj * 4 + 2] += v.z; colsum[gj * 4 + 3] += v.w;
                } else {
                    *reinterpret_cast<float4*>(C + (size_t)row * p.ldc + col) = v;
                }
            }
        }
    }
    if (p.pool_partial) {
        // cross-thread column reduction through shared memory in a fixed (ty ascending) order
        __syncthreads();
        float* red = &As[0][0][0];                       // NY x BN floats  (NY*BN <= 2*BK*(BM+4))
#pragma unroll
        for (int gj = 0; gj < TNG; ++gj)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) red[ty * BN + gj * (BN / TNG) + tx * 4 + jj] = colsum[gj * 4 + jj];
        __syncthreads();
        for (int c = tid; c < BN; c += NT) {
            if (n0 + c < p.N) {
                float s = 0.0f;
                for (int y = 0; y < NY; ++y) s += red[y * BN + c];
                p.pool_partial[((size_t)crop * p.tiles_per_crop + tile_in_crop) * p.N + n0 + c] = s;
            }
        }
    }
}

template <int BM, int BN, int TMG, int TNG>
void launch(const GemmParams& p, int groups, cudaStream_t s)
{
    constexpr int NT = (BM / (4 * TMG)) * (BN / (4 * TNG));
    static_assert((BM / (4 * TMG)) * BN <= 2 * BK * (BM + 4), "pool scratch must fit in As");
    const int mt = p.pool_partial ? (p.M / p.rows_per_crop) * p.tiles_per_crop : (p.M + BM - 1) / BM;
    dim3 grid((p.N + BN - 1) / BN, mt, groups);
    sgemm_kernel<BM, BN, TMG, TNG><<<grid, NT, 0, s>>>(p);
}

}  // namespace

extern "C" int df_gemm_rows_per_pool_tile(void) { return 128; }

static int gemm_fp32_impl(const float* A, int lda, const float* W, int ldw, const float* bias,
                          int bias_crop_stride, float* C, int ldc, int M, int N, int K, int relu,
                          int rows_per_crop, int groups, long long a_group_stride, long long w_group_stride,
                          long long bias_group_stride, long long c_group_stride, float* pool_partial,
                          const float* relu_mask, int accumulate, void* stream);

extern "C" int df_gemm_fp32(const float* A, int lda, const float* W, int ldw, const float* bias,
                            int bias_crop_stride, float* C, int ldc, int M, int N, int K, int relu,
                            int rows_per_crop, int groups, long long a_group_stride, long long w_group_stride,
                            long long bias_group_stride, long long c_group_stride, float* pool_partial,
                            void* stream)
{
    return gemm_fp32_impl(A, lda, W, ldw, bias, bias_crop_stride, C, ldc, M, N, K, relu, rows_per_crop, groups,
                          a_group_stride, w_group_stride, bias_group_stride, c_group_stride, pool_partial, nullptr, 0, stream);
}

// Data-gradient form: C (+)= (A . W^T) (*) [relu_mask > 0], with A = dY and W = the TRANSPOSED layer weight.
extern "C" int df_gemm_dgrad_fp32(const float* dY, int ldy, const float* Wt, int ldw, float* dX, int ldx, int M, int N,
                                  int K, int groups, long long dy_group_stride, long long w_group_stride,
                                  long long dx_group_stride, const float* relu_mask, int accumulate, void* stream)
{
    if (!dX) return DF_ERR_ARG;
    return gemm_fp32_impl(dY, ldy, Wt, ldw, nullptr, 0, dX, ldx, M, N, K, 0, 0, groups, dy_group_stride, w_group_stride, 0,
                          dx_group_stride, nullptr, relu_mask, accumulate, stream);
}

static int gemm_fp32_impl(const float* A, int lda, const float* W, int ldw, const float* bias,
                          int bias_crop_stride, float* C, int ldc, int M, int N, int K, int relu,
                          int rows_per_crop, int groups, long long a_group_stride, long long w_group_stride,
                          long long bias_group_stride, long long c_group_stride, float* pool_partial,
                          const float* relu_mask, int accumulate, void* stream)
{
    if (!A || !W || (!C && !pool_partial)) return DF_ERR_ARG;
    if (M <= 0 || N <= 0 || K <= 0 || groups <= 0) return DF_ERR_ARG;
    if (K % BK || lda % 4 || ldw % 4 || N % 4) return DF_ERR_ARG;
    if (C && (ldc % 4 || c_group_stride % 4)) return DF_ERR_ARG;
    if (a_group_stride % 4 || w_group_stride % 4 || bias_group_stride % 4 || bias_crop_stride % 4) return DF_ERR_ARG;
    if ((bias_crop_stride || pool_partial) && rows_per_crop <= 0) return DF_ERR_ARG;
    if (pool_partial && M % rows_per_crop) return DF_ERR_ARG;
    GemmParams p;
    p.A = A; p.lda = lda; p.W = W; p.ldw = ldw; p.bias = bias; p.bias_crop_stride = bias_crop_stride;
    p.C = pool_partial ? nullptr : C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.relu = relu;
    p.rows_per_crop = rows_per_crop > 0 ? rows_per_crop : M;
    p.a_gs = a_group_stride; p.w_gs = w_group_stride; p.bias_gs = bias_group_stride; p.c_gs = c_group_stride;
    p.pool_partial = pool_partial;
    p.tiles_per_crop = pool_partial ? (p.rows_per_crop + 127) / 128 : 0;
    p.relu_mask = relu_mask; p.accumulate = accumulate;
    if (((uintptr_t)relu_mask & 15) || (pool_partial && (relu_mask || accumulate))) return DF_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (pool_partial) {
        if (N % 128 == 0) launch<128, 128, 2, 2>(p, groups, s);
        else launch<128, 64, 2, 1>(p, groups, s);
    } else if (M <= 64) {
        launch<32, 128, 1, 1>(p, groups, s);
    } else if (N % 128 == 0) {
        launch<128, 128, 2, 2>(p, groups, s);
    } else {
        launch<128, 64, 2, 1>(p, groups, s);
    }
    DF_RETURN_LAST_ERROR();
}
