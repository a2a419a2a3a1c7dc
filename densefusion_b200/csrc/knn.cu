// K4 -- brute-force nearest neighbours, drop-in for lib/knn (reference:
// lib/knn/src/knn_cuda_kernel.cu:31-170 + the C glue lib/knn/src/knn_pytorch.c:6-48).
//
// The reference materialises the full R x Q distance matrix in HBM (8*R*Q bytes of traffic) and
// then scans each column.  Here the reference points are staged once per CTA in shared memory as
// float4 (one broadcast LDS.128 per reference point serves every query of the warp), each thread
// keeps KNN_QPT queries and their running minimum in registers, and nothing but the inputs and the
// int64 indices ever touches HBM (12*R + 20*Q bytes).  Distances use the reference's FMA chain and
// the strict '<' update, so indices are bit-identical including ties (lowest index) and NaN rows.
#include "df_common.cuh"
#include "../../include/densefusion_b200.h"

namespace {

constexpr int KNN_THREADS = 256;
constexpr int KNN_TILE = 2048;   // reference points per shared-memory tile (32 KB as float4)
constexpr long KNN_WARP_MAX_QUERIES = 148L * 8 * 8 * 2;   // up to two waves of warps (8 CTAs x 8 warps per SM): 18 944 queries

// ---- D = 3, k = 1 : the path's own shape (lib/loss.py:42-47, tools/eval_linemod.py:124-128) ----
template <int QPT>
__global__ void __launch_bounds__(KNN_THREADS)
knn1_d3_kernel(const float* __restrict__ ref, const float* __restrict__ query, int64_t* __restrict__ ind,
               int R, int Q)
{
    __shared__ float4 s_ref[KNN_TILE];
    const int b = blockIdx.y;
    ref += (size_t)b * 3 * R;
    query += (size_t)b * 3 * Q;
    ind += (size_t)b * Q;

    const int q0 = blockIdx.x * (KNN_THREADS * QPT) + threadIdx.x;
    float qx[QPT], qy[QPT], qz[QPT], best[QPT];
    int chunk[QPT];
    // row 0 seeds the minimum exactly like `max_dist = p_dist[0]` (knn_cuda_kernel.cu:122)
    const float r0x = __ldg(ref), r0y = __ldg(ref + R), r0z = __ldg(ref + 2 * (size_t)R);
#pragma unroll
    for (int i = 0; i < QPT; ++i) {
        const int q = q0 + i * KNN_THREADS;
        const bool ok = q < Q;
        qx[i] = ok ? __ldg(query + q) : 0.0f;
        qy[i] = ok ? __ldg(query + Q + q) : 0.0f;
        qz[i] = ok ? __ldg(query + 2 * (size_t)Q + q) : 0.0f;
        best[i] = df::ref_ssd3(r0x, r0y, r0z, qx[i], qy[i], qz[i]);
        chunk[i] = 0;
    }
    for (int base = 0; base < R; base += KNN_TILE) {
        const int n = min(KNN_TILE, R - base);
        const int n_pad = (n + df::NN_CHUNK - 1) / df::NN_CHUNK * df::NN_CHUNK;
        __syncthreads();
        for (int r = threadIdx.x; r < n_pad; r += KNN_THREADS)
            s_ref[r] = r < n ? make_float4(__ldg(ref + base + r), __ldg(ref + R + base + r),
                                           __ldg(ref + 2 * (size_t)R + base + r), 0.0f)
                             : make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.0f);
        __syncthreads();
        df::nn_scan_tile<QPT>(s_ref, n_pad, base, qx, qy, qz, best, chunk);
    }
#pragma unroll
    for (int i = 0; i < QPT; ++i) {
        const int q = q0 + i * KNN_THREADS;
        if (q < Q)   // 1-based (knn_cuda_kernel.cu:123,163)
            ind[q] = (int64_t)df::nn_resolve<false>(ref, R, chunk[i], best[i], qx[i], qy[i], qz[i]) + 1;
    }
}

// ---- D = 3, k = 1, FEW queries (lib/loss_refiner.py:40-46 and tools/eval_linemod.py:124-128: Q = R = 500 ... 2600) ----
// One query per thread leaves 2-11 CTAs on 148 SMs and every thread a serial scan of all R references.  Here a WARP owns a
// query: lane l scans references l, l+32, ... (coalesced, dim-major) keeping its (min, first index) under the strict '<', then
// the 32 partial results are reduced lexicographically on (distance, index) by shuffles, so the LOWEST index wins a tie exactly
// as in the reference's sequential scan (knn_cuda_kernel.cu:150-167).  NaN: a NaN candidate never satisfies '<' (skipped, as in
// the reference); a NaN distance to row 0 -- the reference's seed, knn_cuda_kernel.cu:122 -- never loses, so the answer is row 0;
// an all-(+inf) column also keeps the seed.
__global__ void __launch_bounds__(KNN_THREADS)
knn1_d3_warp_kernel(const float* __restrict__ ref, const float* __restrict__ query, int64_t* __restrict__ ind, int R, int Q)
{
    const int b = blockIdx.y;
    ref += (size_t)b * 3 * R;
    query += (size_t)b * 3 * Q;
    ind += (size_t)b * Q;
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (KNN_THREADS / 32) + (threadIdx.x >> 5);
    if (q >= Q) return;                                         // whole warps leave together
    const float qx = __ldg(query + q), qy = __ldg(query + Q + q), qz = __ldg(query + 2 * (size_t)Q + q);
    float best = CUDART_INF_F;
    int bi = 0x7fffffff;
    for (int r = lane; r < R; r += 32) {
        const float d = df::ref_ssd3(__ldg(ref + r), __ldg(ref + R + r), __ldg(ref + 2 * (size_t)R + r), qx, qy, qz);
        if (d < best) { best = d; bi = r; }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (od < best || (od == best && oi < bi)) { best = od; bi = oi; }
    }
    if (lane == 0) {
        const float d0 = df::ref_ssd3(__ldg(ref), __ldg(ref + R), __ldg(ref + 2 * (size_t)R), qx, qy, qz);
        const int idx = (d0 != d0 || bi == 0x7fffffff) ? 0 : bi;
        ind[q] = (int64_t)idx + 1;                              // 1-based (knn_cuda_kernel.cu:123,163)
    }
}

// ---- general (D, k): the k-slot insertion of the reference's cuInsertionSort restated with register arrays, distances on
// the fly (needed to reproduce the k > 1 ordering quirks bit for bit; not on the pose path) ----------
constexpr int KNN_KMAX = 64;

__global__ void __launch_bounds__(128)
knn_general_kernel(const float* __restrict__ ref, const float* __restrict__ query, int64_t* __restrict__ ind,
                   int R, int Q, int D, int k)
{
    const int b = blockIdx.y;
    ref += (size_t)b * D * R;
    query += (size_t)b * D * Q;
    ind += (size_t)b * k * Q;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;

    float dist[KNN_KMAX];
    int id[KNN_KMAX];
    auto ssd = [&](int r) {
        float s = 0.0f;
        for (int d = 0; d < D; ++d) {
            const float t = __fsub_rn(__ldg(ref + (size_t)d * R + r), __ldg(query + (size_t)d * Q + q));
            s = __fmaf_rn(t, t, s);
        }
        return s;
    };
    // phase 1: order the first k rows (knn_cuda_kernel.cu:126-146)
    dist[0] = ssd(0);
    id[0] = 1;
    float max_dist = dist[0];
    for (int l = 1; l < k; ++l) {
        const float curr = ssd(l);
        dist[l] = curr;
        if (curr < max_dist) {
            int i = l - 1;
            for (int a = 0; a < l - 1; ++a)
                if (dist[a] > curr) { i = a; break; }
            for (int j = l; j > i; --j) { dist[j] = dist[j - 1]; id[j] = id[j - 1]; }
            dist[i] = curr;
            id[i] = l + 1;
        } else {
            id[l] = l + 1;
        }
        max_dist = dist[l];
    }
    // phase 2: stream the remaining rows through the window (knn_cuda_kernel.cu:149-168)
    for (int l = k; l < R; ++l) {
        const float curr = ssd(l);
        if (curr < max_dist) {
            int i = k - 1;
            for (int a = 0; a < k - 1; ++a)
                if (dist[a] > curr) { i = a; break; }
            for (int j = k - 1; j > i; --j) { dist[j] = dist[j - 1]; id[j] = id[j - 1]; }
            dist[i] = curr;
            id[i] = l + 1;
            max_dist = dist[k - 1];
        }
    }
    for (int j = 0; j < k; ++j) ind[(size_t)j * Q + q] = id[j];
}

}  // namespace

extern "C" int df_knn(const float* ref, const float* query, int64_t* ind, int batch, int dim, int R, int Q,
                      int k, void* stream)
{
    if (!ref || !query || !ind || batch <= 0 || dim <= 0 || R <= 0 || Q < 0 || k <= 0 || k > R) return DF_ERR_ARG;
    if (batch > 65535) return DF_ERR_ARG;
    if (Q == 0) return DF_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (dim == 3 && k == 1) {
        // enough CTAs for >= 2 waves of 148 SMs x 8 resident CTAs when Q allows, else fewer queries/thread
        const long per_wave = 148L * 8 * KNN_THREADS;
        if ((long)Q * batch <= KNN_WARP_MAX_QUERIES) {         // few queries: a warp per query (see knn1_d3_warp_kernel)
            dim3 grid((Q + KNN_THREADS / 32 - 1) / (KNN_THREADS / 32), batch);
            knn1_d3_warp_kernel<<<grid, KNN_THREADS, 0, s>>>(ref, query, ind, R, Q);
        } else if ((long)Q * batch >= 8 * per_wave) {
            dim3 grid((Q + KNN_THREADS * 4 - 1) / (KNN_THREADS * 4), batch);
            knn1_d3_kernel<4><<<grid, KNN_THREADS, 0, s>>>(ref, query, ind, R, Q);
        } else if ((long)Q * batch >= 2 * per_wave) {
            dim3 grid((Q + KNN_THREADS * 2 - 1) / (KNN_THREADS * 2), batch);
            knn1_d3_kernel<2><<<grid, KNN_THREADS, 0, s>>>(ref, query, ind, R, Q);
        } else {
            dim3 grid((Q + KNN_THREADS - 1) / KNN_THREADS, batch);
            knn1_d3_kernel<1><<<grid, KNN_THREADS, 0, s>>>(ref, query, ind, R, Q);
        }
    } else {
        if (k > KNN_KMAX) return DF_ERR_UNSUPPORTED;
        dim3 grid((Q + 127) / 128, batch);
        knn_general_kernel<<<grid, 128, 0, s>>>(ref, query, ind, R, Q, dim, k);
    }
    DF_RETURN_LAST_ERROR();
}
