// Device-side crop builder: what tools/eval_ycb.py:147-190 does per object in numpy --
// mask = (label == item) & (depth != 0) inside the snapped bounding box, the N sampled pixel indices `choose`, the
// back-projected cloud and the normalised colour crop -- for a whole bucket of objects in two launches, so the host
// only ships raw frames.  HBM-bound byte work: one CTA per object streams its box once for the row counts and once
// for the picks; the colour crop is a coalesced gather.
//
// Arithmetic is the reference's, operation by operation in fp32 (numpy keeps float32 through every step):
//   pt2 = depth / cam_scale;  pt0 = (col - cx) * pt2 / fx;  pt1 = (row - cy) * pt2 / fy;  img = (v - mean) / std
// Sampling: with at most N masked pixels the list is wrap-padded exactly like np.pad(..., 'wrap'); with more, the
// reference draws a uniformly random subset (np.random.shuffle) -- here ranks floor((j + u) * count / N), j = 0..N-1,
// with a per-object offset u in [0,1): ascending, duplicate-free, every pixel equally likely, reproducible from `seed`.
#include "df_common.cuh"
#include "../../include/densefusion_b200.h"

namespace {

constexpr int MAX_ROWS = 1024;

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// meta (b,6) int32: frame, item id, rmin, rmax, cmin, cmax
__global__ void __launch_bounds__(256)
crop_select_kernel(const float* __restrict__ depth, const int* __restrict__ label, const int* __restrict__ meta, int H, int W,
                   int N, float cx, float cy, float fx, float fy, float scale, uint32_t seed, int64_t* __restrict__ choose,
                   float* __restrict__ cloud, int* __restrict__ count)
{
    __shared__ int rowpre[MAX_ROWS + 1];
    const int o = blockIdx.x, tid = threadIdx.x;
    const int frame = meta[o * 6], item = meta[o * 6 + 1];
    const int rmin = meta[o * 6 + 2], rmax = meta[o * 6 + 3], cmin = meta[o * 6 + 4], cmax = meta[o * 6 + 5];
    const int rows = rmax - rmin, cols = cmax - cmin;
    const float* d = depth + (size_t)frame * H * W;
    const int* l = label + (size_t)frame * H * W;
    for (int r = tid; r < rows; r += blockDim.x) {
        int c = 0;
        const size_t base = (size_t)(rmin + r) * W;
        for (int x = cmin; x < cmax; ++x) c += (l[base + x] == item && d[base + x] != 0.0f) ? 1 : 0;
        rowpre[r + 1] = c;
    }
    if (tid == 0) rowpre[0] = 0;
    __syncthreads();
    if (tid == 0)
        for (int r = 0; r < rows; ++r) rowpre[r + 1] += rowpre[r];
    __syncthreads();
    const int total = rowpre[rows];
    if (tid == 0) count[o] = total;
    const float u = (float)(hash32(seed ^ (0x9e3779b9u * (uint32_t)(o + 1))) >> 8) * (1.0f / 16777216.0f);
    for (int j = tid; j < N; j += blockDim.x) {
        int64_t ch = 0;
        float p0 = 0.f, p1 = 0.f, p2 = 0.f;
        if (total > 0) {
            long long rank;
            if (total <= N) rank = j % total;
            else {
                rank = (long long)(((double)j + (double)u) * (double)total / (double)N);
                if (rank >= total) rank = total - 1;
            }
            int lo = 0, hi = rows;                    // largest row with rowpre[row] <= rank
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (rowpre[mid] <= rank) lo = mid; else hi = mid;
            }
            int need = (int)(rank - rowpre[lo]);
            const size_t base = (size_t)(rmin + lo) * W;
            int x = cmin;
            for (; x < cmax; ++x) {
                if (l[base + x] == item && d[base + x] != 0.0f) {
                    if (need == 0) break;
                    --need;
                }
            }
            ch = (int64_t)lo * cols + (x - cmin);
            const float dep = d[base + x];
            p2 = __fdiv_rn(dep, scale);
            p0 = __fdiv_rn(__fmul_rn(__fsub_rn((float)x, cx), p2), fx);
            p1 = __fdiv_rn(__fmul_rn(__fsub_rn((float)(rmin + lo), cy), p2), fy);
        }
        choose[(size_t)o * N + j] = ch;
        float* c3 = cloud + ((size_t)o * N + j) * 3;
        c3[0] = p0; c3[1] = p1; c3[2] = p2;
    }
}

// out (b,3,h,w) = (rgb[frame, rmin+y, cmin+x, c] - mean[c]) / std[c]   (HWC uint8 frames, no 1/255 -- as the reference)
__global__ void __launch_bounds__(256)
crop_image_kernel(const uint8_t* __restrict__ rgb, const int* __restrict__ meta, int H, int W, int h, int w, int b,
                  float m0, float m1, float m2, float s0, float s1, float s2, float* __restrict__ out)
{
    const unsigned total = (unsigned)b * 3 * h * w;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int x = i % w;
        unsigned r = i / w;
        const int y = r % h; r /= h;
        const int c = r % 3, o = r / 3;
        const int frame = meta[o * 6], rmin = meta[o * 6 + 2], cmin = meta[o * 6 + 4];
        const float v = (float)rgb[(((size_t)frame * H + rmin + y) * W + cmin + x) * 3 + c];
        const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
        out[i] = __fdiv_rn(__fsub_rn(v, mean), sd);
    }
}

}  // namespace

extern "C" int df_build_crops(const uint8_t* rgb, const float* depth, const int* label, const int* meta, int b, int H, int W,
                              int h, int w, int N, const float* cam /* cx cy fx fy scale */, const float* mean_std /* 6 */,
                              unsigned seed, float* out_img, int64_t* out_choose, float* out_cloud, int* out_count,
                              void* stream)
{
    if (!rgb || !depth || !label || !meta || !cam || !mean_std || !out_img || !out_choose || !out_cloud || !out_count)
        return DF_ERR_ARG;
    if (b <= 0 || H <= 0 || W <= 0 || h <= 0 || w <= 0 || N <= 0 || h > H || w > W || h > MAX_ROWS) return DF_ERR_ARG;
    if ((long long)b * 3 * h * w >= (1LL << 31)) return DF_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    crop_select_kernel<<<b, 256, 0, s>>>(depth, label, meta, H, W, N, cam[0], cam[1], cam[2], cam[3], cam[4], seed, out_choose,
                                         out_cloud, out_count);
    const long long total = (long long)b * 3 * h * w;
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    crop_image_kernel<<<(unsigned)blocks, 256, 0, s>>>(rgb, meta, H, W, h, w, b, mean_std[0], mean_std[1], mean_std[2],
                                                       mean_std[3], mean_std[4], mean_std[5], out_img);
    DF_RETURN_LAST_ERROR();
}
