// K3 -- fused ADD / ADD-S loss, hypothesis scoring, argmax pose selection and refiner-input
// construction (reference: lib/loss.py:13-70 and lib/loss_refiner.py:12-62), forward and backward.
//
// One launch does what the reference spreads over ~35 torch kernels and six (N,M,3) temporaries:
//   CTA (p, b): hypothesis p of crop b.  q/|q| -> R; each thread transforms its model points
//   (kept in registers), for symmetric objects scans the target cloud staged in shared memory with
//   the bit-exact 1-NN of knn.cu, accumulates |pred - tgt| and the two sums the backward needs
//   (sum u and sum u m^T with u the unit residual), block-reduces in a fixed order.
//   The LAST CTA of a crop to finish (atomic ticket) reduces the confidence-weighted loss,
//   takes the first-index argmax of the confidences and writes new_points / new_target.
// Bytes per crop: 34 KB in, 12 KB out (+ 52*P bytes of per-hypothesis state kept for backward);
// nothing of size P x M is ever stored.
#include "df_common.cuh"
#include "../../include/densefusion_b200.h"
#include <math_constants.h>
#include <stdlib.h>
#include <cooperative_groups.h>

namespace {

constexpr int LOSS_THREADS = 128;
constexpr int LOSS_QPT = 4;                    // model points per thread per pass
constexpr int LOSS_TILE = 1024;                // target points per shared-memory tile
constexpr int LOSS_NRED = 13;                  // dis, sum u (3), sum u m^T (9)

struct LossParams {
    const float* pred_r;      // (B,P,4) un-normalised quaternions (w,x,y,z)
    const float* pred_t;      // (B,P,3)
    const float* pred_c;      // (B,P) confidences, or nullptr (refiner loss: P == 1, no confidence)
    const float* target;      // (B,M,3)
    const float* model;       // (B,M,3)
    const float* hyp_points;  // (B,P,3) per-hypothesis anchor points, or nullptr (refiner loss)
    const float* points;      // (B,N,3) cloud to re-express for the next stage
    const int64_t* idx;       // (B) object ids
    unsigned long long sym_mask;
    int allow_sym;            // 0: never take the symmetric branch (Loss(..., refine=True))
    float w;
    int P, M, N;
    float* dis_all;           // (B,P)
    float* sum_u;             // (B,P,3)   (1/M) sum_j u_j
    float* sum_um;            // (B,P,9)   (1/M) sum_j u_j m_j^T
    float* loss;              // (B)
    float* dis_sel;           // (B)
    int64_t* which;           // (B)
    float* new_points;        // (B,N,3)
    float* new_target;        // (B,M,3)
    unsigned int* tickets;    // (B) zero-initialised; reset by the kernel
    float* dbg_pred;          // optional (B,P,M,3)
    int* dbg_nn;              // optional (B,P,M) 0-based nearest target index (-1 when not symmetric)
};

__device__ __forceinline__ void load_rotation(const float* q, float (&R)[9])
{
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    const float n = sqrtf(w * w + x * x + y * y + z * z);      // torch.norm, no eps (lib/loss.py:16)
    df::quat_to_rot(w / n, x / n, y / n, z / n, R);
}

// CL > 1 (the refiner loss, P == 1: lib/loss_refiner.py:12-62 has ONE hypothesis per crop, i.e. one CTA per crop on 148 SMs): a
// thread-block cluster of CL CTAs shares the hypothesis -- CTA `rank` takes the rank-th slice of the model points (each scans the
// whole target cloud for its own queries), leaves its 13 partial sums in its shared memory, and rank 0 adds them in rank order
// through distributed shared memory (fixed order: deterministic) before it continues as the crop's last CTA.
template <int CL>
__global__ void __launch_bounds__(LOSS_THREADS)
loss_forward_kernel(const LossParams a)
{
    __shared__ float4 s_tgt[LOSS_TILE];
    __shared__ float s_red[LOSS_THREADS / 32][LOSS_NRED];
    __shared__ float s_sel[16];
    __shared__ float s_part[LOSS_NRED];
    __shared__ int s_flag;

    const int p = blockIdx.x / CL, b = blockIdx.y, tid = threadIdx.x;
    const int rank = CL > 1 ? (int)cooperative_groups::this_cluster().block_rank() : 0;
    const int P = a.P, M = a.M;
    const int slice = (M + CL - 1) / CL;                   // model points per CTA of the cluster
    const int m_lo = rank * slice, m_hi = min(M, m_lo + slice);
    const float* tgt = a.target + (size_t)b * M * 3;
    const float* mdl = a.model + (size_t)b * M * 3;
    const long long obj = a.idx[b];
    const bool sym = a.allow_sym && obj >= 0 && obj < 64 && ((a.sym_mask >> obj) & 1ull);

    float R[9];
    load_rotation(a.pred_r + ((size_t)b * P + p) * 4, R);
    float off[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        off[c] = a.pred_t[((size_t)b * P + p) * 3 + c];
        if (a.hyp_points) off[c] = a.hyp_points[((size_t)b * P + p) * 3 + c] + off[c];   // points + pred_t
    }

    float acc[LOSS_NRED];
#pragma unroll
    for (int i = 0; i < LOSS_NRED; ++i) acc[i] = 0.0f;

    const float t0x = tgt[0], t0y = tgt[1], t0z = tgt[2];
    for (int j0 = m_lo; j0 < m_hi; j0 += LOSS_THREADS * LOSS_QPT) {
        float mx[LOSS_QPT], my[LOSS_QPT], mz[LOSS_QPT], px[LOSS_QPT], py[LOSS_QPT], pz[LOSS_QPT];
        float best[LOSS_QPT];
        int arg[LOSS_QPT];
#pragma unroll
        for (int i = 0; i < LOSS_QPT; ++i) {
            const int j = j0 + i * LOSS_THREADS + tid;
            const bool ok = j < m_hi;
            mx[i] = ok ? mdl[j * 3 + 0] : 0.0f;
            my[i] = ok ? mdl[j * 3 + 1] : 0.0f;
            mz[i] = ok ? mdl[j * 3 + 2] : 0.0f;
            // pred = model . R^T + (points + pred_t): bmm accumulates over the model coordinate
            px[i] = (mx[i] * R[0] + my[i] * R[1] + mz[i] * R[2]) + off[0];
            py[i] = (mx[i] * R[3] + my[i] * R[4] + mz[i] * R[5]) + off[1];
            pz[i] = (mx[i] * R[6] + my[i] * R[7] + mz[i] * R[8]) + off[2];
            arg[i] = j;
            if (sym) { best[i] = df::ref_ssd3(t0x, t0y, t0z, px[i], py[i], pz[i]); arg[i] = 0; }
        }
        if (sym) {
            for (int base = 0; base < M; base += LOSS_TILE) {
                const int n = min(LOSS_TILE, M - base);
                const int n_pad = (n + df::NN_CHUNK - 1) / df::NN_CHUNK * df::NN_CHUNK;
                __syncthreads();
                for (int r = tid; r < n_pad; r += LOSS_THREADS)
                    s_tgt[r] = r < n ? make_float4(tgt[(base + r) * 3], tgt[(base + r) * 3 + 1], tgt[(base + r) * 3 + 2], 0.0f)
                                     : make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.0f);
                __syncthreads();
                df::nn_scan_tile<LOSS_QPT>(s_tgt, n_pad, base, px, py, pz, best, arg);      // arg = winning chunk
            }
#pragma unroll
            for (int i = 0; i < LOSS_QPT; ++i) arg[i] = df::nn_resolve<true>(tgt, M, arg[i], best[i], px[i], py[i], pz[i]);
        }
#pragma unroll
        for (int i = 0; i < LOSS_QPT; ++i) {
            const int j = j0 + i * LOSS_THREADS + tid;
            if (j < m_hi) {
                const float ex = px[i] - tgt[arg[i] * 3 + 0];
                const float ey = py[i] - tgt[arg[i] * 3 + 1];
                const float ez = pz[i] - tgt[arg[i] * 3 + 2];
                const float d = sqrtf(ex * ex + ey * ey + ez * ez);
                const float inv = d > 0.0f ? 1.0f / d : 0.0f;      // torch.norm backward: 0 at the origin
                const float ux = ex * inv, uy = ey * inv, uz = ez * inv;
                acc[0] += d;
                acc[1] += ux; acc[2] += uy; acc[3] += uz;
                acc[4] += ux * mx[i]; acc[5] += ux * my[i]; acc[6] += ux * mz[i];
                acc[7] += uy * mx[i]; acc[8] += uy * my[i]; acc[9] += uy * mz[i];
                acc[10] += uz * mx[i]; acc[11] += uz * my[i]; acc[12] += uz * mz[i];
                if (a.dbg_pred) {
                    float* o = a.dbg_pred + (((size_t)b * P + p) * M + j) * 3;
                    o[0] = px[i]; o[1] = py[i]; o[2] = pz[i];
                }
                if (a.dbg_nn) a.dbg_nn[((size_t)b * P + p) * M + j] = sym ? arg[i] : -1;
            }
        }
    }

    // fixed-order block reduction: lanes (shuffle tree) -> warps (ascending)
#pragma unroll
    for (int i = 0; i < LOSS_NRED; ++i) acc[i] = df::warp_sum(acc[i]);
    if ((tid & 31) == 0) {
#pragma unroll
        for (int i = 0; i < LOSS_NRED; ++i) s_red[tid >> 5][i] = acc[i];
    }
    __syncthreads();
    if (CL > 1) {
        namespace cg = cooperative_groups;
        cg::cluster_group cluster = cg::this_cluster();
        if (tid < LOSS_NRED) {
            float v = 0.0f;
            for (int wv = 0; wv < LOSS_THREADS / 32; ++wv) v += s_red[wv][tid];
            s_part[tid] = v;
        }
        cluster.sync();                                    // every CTA's partial sums are in its shared memory
        if (rank == 0 && tid < LOSS_NRED) {
            float v = s_part[tid];
            for (int r = 1; r < CL; ++r) v += *cluster.map_shared_rank(&s_part[tid], r);
            s_red[0][tid] = v;
            for (int wv = 1; wv < LOSS_THREADS / 32; ++wv) s_red[wv][tid] = 0.0f;
        }
        cluster.sync();                                    // the remote reads are done: the other CTAs may leave
        if (rank != 0) return;
        __syncthreads();
    }
    if (tid == 0) {
        float tot[LOSS_NRED];
#pragma unroll
        for (int i = 0; i < LOSS_NRED; ++i) {
            tot[i] = 0.0f;
            for (int wv = 0; wv < LOSS_THREADS / 32; ++wv) tot[i] += s_red[wv][i];
            tot[i] /= (float)M;
        }
        const size_t hp = (size_t)b * P + p;
        a.dis_all[hp] = tot[0];
        for (int i = 0; i < 3; ++i) a.sum_u[hp * 3 + i] = tot[1 + i];
        for (int i = 0; i < 9; ++i) a.sum_um[hp * 9 + i] = tot[4 + i];
        __threadfence();
        const unsigned int ticket = atomicAdd(a.tickets + b, 1u);
        s_flag = (ticket == (unsigned int)(P - 1));
    }
    __syncthreads();
    if (!s_flag) return;
    __threadfence();

    // ---------------- last CTA of crop b: loss, argmax, next-stage inputs ----------------
    const float* dis = a.dis_all + (size_t)b * P;
    float lsum = 0.0f, cbest = -CUDART_INF_F;
    int kbest = 0x7fffffff;
    if (a.pred_c) {
        const float* conf = a.pred_c + (size_t)b * P;
        for (int h = tid; h < P; h += LOSS_THREADS) {
            const float c = conf[h];
            lsum += __ldcg(dis + h) * c - a.w * logf(c);
            if (c > cbest) { cbest = c; kbest = h; }       // ascending h: first maximum kept
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
            const float oc = __shfl_xor_sync(0xffffffffu, cbest, o);
            const int ok = __shfl_xor_sync(0xffffffffu, kbest, o);
            if (oc > cbest || (oc == cbest && ok < kbest)) { cbest = oc; kbest = ok; }
        }
        if ((tid & 31) == 0) { s_red[tid >> 5][0] = lsum; s_red[tid >> 5][1] = cbest; s_red[tid >> 5][2] = __int_as_float(kbest); }
        __syncthreads();
        if (tid == 0) {
            float tot = 0.0f, cb = -CUDART_INF_F;
            int kb = 0x7fffffff;
            for (int wv = 0; wv < LOSS_THREADS / 32; ++wv) {
                tot += s_red[wv][0];
                const float oc = s_red[wv][1];
                const int ok = __float_as_int(s_red[wv][2]);
                if (oc > cb || (oc == cb && ok < kb)) { cb = oc; kb = ok; }
            }
            if (kb == 0x7fffffff) kb = 0;                   // all-NaN confidences: degenerate, pick 0
            s_sel[0] = tot / (float)P;
            s_sel[1] = __int_as_float(kb);
        }
        __syncthreads();
        kbest = __float_as_int(s_sel[1]);
        lsum = s_sel[0];
    } else {
        kbest = 0;
        lsum = __ldcg(dis);
    }
    const size_t hk = (size_t)b * P + kbest;
    if (tid == 0) {
        a.loss[b] = lsum;
        a.dis_sel[b] = __ldcg(dis + kbest);
        a.which[b] = kbest;
        a.tickets[b] = 0u;
    }
    float Rk[9], tk[3];
    load_rotation(a.pred_r + hk * 4, Rk);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        tk[c] = a.pred_t[hk * 3 + c];
        if (a.hyp_points) tk[c] = tk[c] + a.hyp_points[hk * 3 + c];       // ori_t[k] + points[k]
    }
    // new = (x - t) . R_k   (bmm with ori_base, lib/loss.py:61-67)
    for (int pass = 0; pass < 2; ++pass) {
        const float* src = pass == 0 ? a.points + (size_t)b * a.N * 3 : tgt;
        float* dst = pass == 0 ? a.new_points + (size_t)b * a.N * 3 : a.new_target + (size_t)b * M * 3;
        const int cnt = pass == 0 ? a.N : M;
        if (!src || !dst) continue;
        for (int n = tid; n < cnt; n += LOSS_THREADS) {
            const float vx = src[n * 3] - tk[0], vy = src[n * 3 + 1] - tk[1], vz = src[n * 3 + 2] - tk[2];
            dst[n * 3 + 0] = vx * Rk[0] + vy * Rk[3] + vz * Rk[6];
            dst[n * 3 + 1] = vx * Rk[1] + vy * Rk[4] + vz * Rk[7];
            dst[n * 3 + 2] = vx * Rk[2] + vy * Rk[5] + vz * Rk[8];
        }
    }
}

// Backward: one thread per hypothesis.  Upstream gradients: g_loss (B) on the per-crop loss and
// g_dis (B) on the returned dis at the selected hypothesis (either may be nullptr == zero).
struct LossBwdParams {
    const float* pred_r; const float* pred_c;
    const float* dis_all; const float* sum_u; const float* sum_um;
    const int64_t* which;
    const float* g_loss; const float* g_dis;
    float w; int P; int B;
    float* g_pred_r; float* g_pred_t; float* g_pred_c;   // (B,P,4) (B,P,3) (B,P)
};

__global__ void loss_backward_kernel(const LossBwdParams a)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= a.B * a.P) return;
    const int b = h / a.P, p = h - b * a.P;
    const float gl = a.g_loss ? a.g_loss[b] : 0.0f;
    const float gsel = (a.g_dis && a.which[b] == p) ? a.g_dis[b] : 0.0f;
    float gd;
    if (a.pred_c) {
        const float c = a.pred_c[h];
        gd = gl * c / (float)a.P + gsel;
        if (a.g_pred_c) a.g_pred_c[h] = gl * (a.dis_all[h] - a.w / c) / (float)a.P;
    } else {
        gd = gl + gsel;          // refiner loss: the loss IS the distance
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) a.g_pred_t[(size_t)h * 3 + i] = gd * a.sum_u[(size_t)h * 3 + i];
    float G[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) G[i] = gd * a.sum_um[(size_t)h * 9 + i];
    const float* q = a.pred_r + (size_t)h * 4;
    const float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const float w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
    // dL/dq_hat from dL/dR (R as in df::quat_to_rot)
    const float gw = 2.0f * (-z * G[1] + y * G[2] + z * G[3] - x * G[5] - y * G[6] + x * G[7]);
    const float gx = 2.0f * (y * G[1] + z * G[2] + y * G[3] - 2.0f * x * G[4] - w * G[5] + z * G[6] + w * G[7] - 2.0f * x * G[8]);
    const float gy = 2.0f * (-2.0f * y * G[0] + x * G[1] + w * G[2] + x * G[3] + z * G[5] - w * G[6] + z * G[7] - 2.0f * y * G[8]);
    const float gz = 2.0f * (-2.0f * z * G[0] - w * G[1] + x * G[2] + w * G[3] - 2.0f * z * G[4] + y * G[5] + x * G[6] + y * G[7]);
    // through q_hat = q / |q|
    const float dot = w * gw + x * gx + y * gy + z * gz;
    float* o = a.g_pred_r + (size_t)h * 4;
    o[0] = (gw - w * dot) / n;
    o[1] = (gx - x * dot) / n;
    o[2] = (gy - y * dot) / n;
    o[3] = (gz - z * dot) / n;
}

}  // namespace

extern "C" int df_loss_forward(const float* pred_r, const float* pred_t, const float* pred_c, const float* target,
                               const float* model_points, const float* hyp_points, const float* points,
                               const int64_t* idx, unsigned long long sym_mask, int allow_sym, float w,
                               int B, int P, int M, int N,
                               float* dis_all, float* sum_u, float* sum_um, float* loss, float* dis_sel,
                               int64_t* which, float* new_points, float* new_target, unsigned int* tickets,
                               float* dbg_pred, int* dbg_nn, void* stream)
{
    if (!pred_r || !pred_t || !target || !model_points || !idx || !dis_all || !sum_u || !sum_um || !loss ||
        !dis_sel || !which || !tickets)
        return DF_ERR_ARG;
    if (B <= 0 || B > 65535 || P <= 0 || M <= 0 || N < 0) return DF_ERR_ARG;
    if (!pred_c && P != 1) return DF_ERR_ARG;
    LossParams a;
    a.pred_r = pred_r; a.pred_t = pred_t; a.pred_c = pred_c; a.target = target; a.model = model_points;
    a.hyp_points = hyp_points; a.points = points; a.idx = idx; a.sym_mask = sym_mask; a.allow_sym = allow_sym;
    a.w = w; a.P = P; a.M = M; a.N = N; a.dis_all = dis_all; a.sum_u = sum_u; a.sum_um = sum_um; a.loss = loss;
    a.dis_sel = dis_sel; a.which = which; a.new_points = new_points; a.new_target = new_target;
    a.tickets = tickets; a.dbg_pred = dbg_pred; a.dbg_nn = dbg_nn;
    // One hypothesis per crop (the refiner loss) and few crops (training runs it at batch 1, tools/train.py:159): 8 CTAs of a cluster
    // share the hypothesis.  Measured (profiles/r2_s4_loss_refine_probe.jsonl, us per call inside a CUDA graph): B = 1, M = 500
    // 27.8 -> 21.0; B = 1, M = 2600 563 -> 70.7; B = 8 27.9 -> 20.6; with 256 crops the 2048 CTAs each scanning the whole target cloud for
    // 63 queries lose (33.6 -> 145), hence the limit on B.  The sums of the slices are added in a different order than one CTA adds
    // its warps: results differ in the last bits between the two forms, each form is deterministic.  DF_LOSS_CLUSTER=0: never.
    static const int use_cluster = getenv("DF_LOSS_CLUSTER") ? atoi(getenv("DF_LOSS_CLUSTER")) : 1;
    if (P == 1 && M >= 64 && B <= 16 && use_cluster) {
        constexpr int CL = 8;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(CL, B); cfg.blockDim = dim3(LOSS_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, loss_forward_kernel<CL>, a);
        if (e != cudaSuccess) return (int)e;
        DF_RETURN_LAST_ERROR();
    }
    loss_forward_kernel<1><<<dim3(P, B), LOSS_THREADS, 0, (cudaStream_t)stream>>>(a);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_loss_backward(const float* pred_r, const float* pred_c, const float* dis_all, const float* sum_u,
                                const float* sum_um, const int64_t* which, const float* g_loss, const float* g_dis,
                                float w, int B, int P, float* g_pred_r, float* g_pred_t, float* g_pred_c,
                                void* stream)
{
    if (!pred_r || !dis_all || !sum_u || !sum_um || !which || !g_pred_r || !g_pred_t) return DF_ERR_ARG;
    if (B <= 0 || P <= 0) return DF_ERR_ARG;
    LossBwdParams a;
    a.pred_r = pred_r; a.pred_c = pred_c; a.dis_all = dis_all; a.sum_u = sum_u; a.sum_um = sum_um;
    a.which = which; a.g_loss = g_loss; a.g_dis = g_dis; a.w = w; a.P = P; a.B = B;
    a.g_pred_r = g_pred_r; a.g_pred_t = g_pred_t; a.g_pred_c = g_pred_c;
    const int total = B * P;
    loss_backward_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
    DF_RETURN_LAST_ERROR();
}
