/*
 * oracle/knn_ref.c -- TEST INFRASTRUCTURE ONLY (never linked or imported by the product).
 *
 * Plain-C restatement of the reference's brute-force k-NN so that index parity can be
 * checked bit-for-bit without a GPU.  It follows, statement by statement, the arithmetic
 * and the comparison order of the two reference kernels:
 *
 *   distance  : /root/reference/lib/knn/src/knn_cuda_kernel.cu:31-95  (cuComputeDistanceGlobal)
 *               ssd starts at 0 and accumulates tmp*tmp over the point dimension in row order;
 *               nvcc contracts `ssd += tmp*tmp` into one FFMA (SASS probe),
 *               hence fmaf() below.  Rows >= dim are zero padded by the kernel and add nothing.
 *   selection : /root/reference/lib/knn/src/knn_cuda_kernel.cu:107-170 (cuInsertionSort)
 *               running k-smallest with strict '<' (lowest index wins ties, NaN never inserts),
 *               indices are 1-based.
 *   layout    : ref is (dim, R), query is (dim, Q), both dim-major, as handed over by
 *               /root/reference/lib/knn/src/knn_pytorch.c:22-36; ind is (k, Q) int64.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off so only the explicit fmaf fuses).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* distance of one (ref r, query q) pair, reference rounding order */
static inline float pair_ssd(const float *ref, int R, const float *query, int Q, int dim, int r, int q)
{
    float ssd = 0.0f;
    for (int d = 0; d < dim; ++d) {
        float tmp = ref[(size_t)d * R + r] - query[(size_t)d * Q + q];
        ssd = fmaf(tmp, tmp, ssd);
    }
    return ssd;
}

/* The reference sorts one column of the (R x Q) distance matrix in place.  `col` is that
 * column (length R), `ind` receives k 1-based indices.  Same two phases as the kernel. */
static void insertion_select(float *col, int64_t *ind, int R, int k)
{
    float max_dist = col[0];
    ind[0] = 1;
    /* phase 1: order the first k entries */
    for (int l = 1; l < k; ++l) {
        float curr = col[l];
        if (curr < max_dist) {
            int i = l - 1;
            for (int a = 0; a < l - 1; ++a) {
                if (col[a] > curr) { i = a; break; }
            }
            for (int j = l; j > i; --j) { col[j] = col[j - 1]; ind[j] = ind[j - 1]; }
            col[i] = curr;
            ind[i] = l + 1;
        } else {
            ind[l] = l + 1;
        }
        max_dist = col[l];
    }
    /* phase 2: stream the remaining rows through the k-slot window */
    for (int l = k; l < R; ++l) {
        float curr = col[l];
        if (curr < max_dist) {
            int i = k - 1;
            for (int a = 0; a < k - 1; ++a) {
                if (col[a] > curr) { i = a; break; }
            }
            for (int j = k - 1; j > i; --j) { col[j] = col[j - 1]; ind[j] = ind[j - 1]; }
            col[i] = curr;
            ind[i] = l + 1;
            max_dist = col[k - 1];
        }
    }
}

/* ref (dim,R)  query (dim,Q)  ->  ind (k,Q) int64 1-based.  Returns 0, or -1 on bad args. */
int df_oracle_knn(const float *ref, int R, const float *query, int Q, int dim, int k, int64_t *ind)
{
    if (!ref || !query || !ind || R <= 0 || Q < 0 || dim <= 0 || k <= 0 || k > R) return -1;
    int fail = 0;
#pragma omp parallel
    {
        float *col = (float *)malloc((size_t)R * sizeof(float));
        int64_t *sel = (int64_t *)malloc((size_t)(k > R ? k : R) * sizeof(int64_t));
        if (!col || !sel) {
#pragma omp atomic write
            fail = 1;
        } else {
#pragma omp for schedule(static)
            for (int q = 0; q < Q; ++q) {
                for (int r = 0; r < R; ++r) col[r] = pair_ssd(ref, R, query, Q, dim, r, q);
                insertion_select(col, sel, R, k);
                for (int j = 0; j < k; ++j) ind[(size_t)j * Q + q] = sel[j];
            }
        }
        free(col);
        free(sel);
    }
    return fail ? -1 : 0;
}

/* k = 1, dim = 3 fast form used for the large ADD-S cases (Q = num_p * num_pt_mesh).
 * Identical result to df_oracle_knn(..., dim=3, k=1, ...): first row seeds the minimum,
 * later rows replace it only on strict '<'. */
int df_oracle_knn1_d3(const float *ref, int R, const float *query, int Q, int64_t *ind)
{
    if (!ref || !query || !ind || R <= 0 || Q < 0) return -1;
    const float *rx = ref, *ry = ref + R, *rz = ref + 2 * (size_t)R;
    const float *qx = query, *qy = query + Q, *qz = query + 2 * (size_t)Q;
#pragma omp parallel for schedule(static)
    for (int q = 0; q < Q; ++q) {
        float x = qx[q], y = qy[q], z = qz[q];
        float best;
        int64_t arg = 1;
        {
            float dx = rx[0] - x, dy = ry[0] - y, dz = rz[0] - z;
            best = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, 0.0f)));
        }
        for (int r = 1; r < R; ++r) {
            float dx = rx[r] - x, dy = ry[r] - y, dz = rz[r] - z;
            float d = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, 0.0f)));
            if (d < best) { best = d; arg = r + 1; }
        }
        ind[q] = arg;
    }
    return 0;
}

int df_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
