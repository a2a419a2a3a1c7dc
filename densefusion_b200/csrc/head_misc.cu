// Bandwidth-bound pieces of the dense-fusion head around the GEMMs (reference lib/network.py):
//   df_gather_embedding : emb[b,c,n] = feat[b,c,choose[b,n]]                       (:98-102)
//   df_xyz_conv         : relu(conv1(x)), the K=3 layer, written into its concat slot  (:54, :152)
//   df_pool_finish      : AvgPool1d over the points from per-tile column sums        (:65, :165)
//   df_select_out       : conv4_{r,t,c} / conv3_{r,t} for the SELECTED object only, sigmoid on c,
//                         point-major stores == the reference's index_select + transpose (:118-130,
//                         :198-204)
// All are coalesced, vectorised where alignment allows, and each output element is written once.
#include "df_common.cuh"
#include "../../include/densefusion_b200.h"

namespace {

// one warp per point: lane == channel (32 channels)
__global__ void __launch_bounds__(256)
gather_embedding_kernel(const float* __restrict__ feat, const int64_t* __restrict__ choose,
                        float* __restrict__ emb_pm, float* __restrict__ emb_cm,
                        long long sb, long long sc, long long sp, int B, int N, int HW)
{
    const int lane = threadIdx.x & 31;
    const long long pt = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pt >= (long long)B * N) return;
    const int b = (int)(pt / N), n = (int)(pt - (long long)b * N);
    long long pix = choose[pt];
    pix = pix < 0 ? 0 : (pix >= HW ? HW - 1 : pix);     // torch.gather would raise; clamp instead of faulting
    const float v = __ldg(feat + b * sb + lane * sc + pix * sp);
    if (emb_pm) emb_pm[pt * 32 + lane] = v;
    if (emb_cm) emb_cm[((long long)b * 32 + lane) * N + n] = v;
}

// x (rows,3) -> out[row, 0:64] (leading dimension ldo): 16 threads per point, 4 channels each
__global__ void __launch_bounds__(256)
xyz_conv_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                float* __restrict__ out, int ldo, long long rows)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long row = t >> 4;
    const int c0 = (int)(t & 15) * 4;
    if (row >= rows) return;
    const float px = __ldg(x + row * 3), py = __ldg(x + row * 3 + 1), pz = __ldg(x + row * 3 + 2);
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + i;
        float s = fmaf(__ldg(W + c * 3 + 2), pz, fmaf(__ldg(W + c * 3 + 1), py, __ldg(W + c * 3) * px));
        o[i] = fmaxf(s + __ldg(bias + c), 0.0f);
    }
    *reinterpret_cast<float4*>(out + row * ldo + c0) = make_float4(o[0], o[1], o[2], o[3]);
}

__global__ void pool_finish_kernel(const float* __restrict__ partial, float* __restrict__ g, int tiles, int C,
                                   int rows_per_crop, long long total)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long crop = i / C;
    const int c = (int)(i - crop * C);
    float s = 0.0f;
    for (int t = 0; t < tiles; ++t) s += partial[(crop * tiles + t) * C + c];
    g[i] = s / (float)rows_per_crop;
}

// One warp per row.  h: (rows, ldh) with branch slices of 128 channels at slice*128.
// Branch r: 4 outputs, t: 3, c: 1 (optional).  Weights (num_obj*width, 128) row-major.
__global__ void __launch_bounds__(256)
select_out_kernel(const float* __restrict__ h, int ldh, const float* __restrict__ Wr, const float* __restrict__ br,
                  const float* __restrict__ Wt, const float* __restrict__ bt, const float* __restrict__ Wc,
                  const float* __restrict__ bc, const int64_t* __restrict__ obj, int rows_per_crop,
                  int num_obj, long long rows, float* __restrict__ out_r, float* __restrict__ out_t,
                  float* __restrict__ out_c)
{
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    long long o = obj[row / rows_per_crop];
    o = o < 0 ? 0 : (o >= num_obj ? num_obj - 1 : o);
    const float* hr = h + row * ldh;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
    {
        const float4 v = __ldg(reinterpret_cast<const float4*>(hr + lane * 4));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 wv = __ldg(reinterpret_cast<const float4*>(Wr + (o * 4 + i) * 128 + lane * 4));
            acc[i] = v.x * wv.x + v.y * wv.y + v.z * wv.z + v.w * wv.w;
        }
    }
    {
        const float4 v = __ldg(reinterpret_cast<const float4*>(hr + 128 + lane * 4));
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float4 wv = __ldg(reinterpret_cast<const float4*>(Wt + (o * 3 + i) * 128 + lane * 4));
            acc[4 + i] = v.x * wv.x + v.y * wv.y + v.z * wv.z + v.w * wv.w;
        }
    }
    if (Wc) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(hr + 256 + lane * 4));
        const float4 wv = __ldg(reinterpret_cast<const float4*>(Wc + o * 128 + lane * 4));
        acc[7] = v.x * wv.x + v.y * wv.y + v.z * wv.z + v.w * wv.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = df::warp_sum(acc[i]);
    if (lane < 4) out_r[row * 4 + lane] = acc[lane] + __ldg(br + o * 4 + lane);
    else if (lane < 7) out_t[row * 3 + (lane - 4)] = acc[lane] + __ldg(bt + o * 3 + (lane - 4));
    else if (lane == 7 && Wc) {
        const float z = acc[7] + __ldg(bc + o);
        out_c[row] = 1.0f / (1.0f + expf(-z));
    }
}

}  // namespace

extern "C" int df_gather_embedding(const float* feat, const int64_t* choose, float* emb_pm, float* emb_cm,
                                   long long stride_b, long long stride_c, long long stride_pix, int B, int N,
                                   int HW, void* stream)
{
    if (!feat || !choose || (!emb_pm && !emb_cm) || B <= 0 || N <= 0 || HW <= 0) return DF_ERR_ARG;
    const long long pts = (long long)B * N;
    gather_embedding_kernel<<<(unsigned)((pts + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        feat, choose, emb_pm, emb_cm, stride_b, stride_c, stride_pix, B, N, HW);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_xyz_conv(const float* x, const float* W, const float* bias, float* out, int ldo, long long rows,
                           void* stream)
{
    if (!x || !W || !bias || !out || rows <= 0 || ldo < 64 || ldo % 4) return DF_ERR_ARG;
    const long long threads = rows * 16;
    xyz_conv_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, W, bias, out, ldo, rows);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_pool_finish(const float* partial, float* g, int crops, int tiles, int channels, int rows_per_crop,
                              void* stream)
{
    if (!partial || !g || crops <= 0 || tiles <= 0 || channels <= 0 || rows_per_crop <= 0) return DF_ERR_ARG;
    const long long total = (long long)crops * channels;
    pool_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(partial, g, tiles, channels,
                                                                                         rows_per_crop, total);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_select_out(const float* h, int ldh, const float* Wr, const float* br, const float* Wt,
                             const float* bt, const float* Wc, const float* bc, const int64_t* obj,
                             int rows_per_crop, int num_obj, long long rows, float* out_r, float* out_t,
                             float* out_c, void* stream)
{
    if (!h || !Wr || !br || !Wt || !bt || !obj || !out_r || !out_t || rows <= 0 || rows_per_crop <= 0 || num_obj <= 0)
        return DF_ERR_ARG;
    if ((Wc && (!bc || !out_c)) || ldh % 4 || ldh < (Wc ? 384 : 256)) return DF_ERR_ARG;
    select_out_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        h, ldh, Wr, br, Wt, bt, Wc, bc, obj, rows_per_crop, num_obj, rows, out_r, out_t, out_c);
    DF_RETURN_LAST_ERROR();
}
