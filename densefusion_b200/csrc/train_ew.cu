// Element-wise / pooling kernels of the colour encoder's TRAINING graph (reference: the autograd of lib/extractors.py:78-124 and
// lib/pspnet.py:7-77) -- what round 1 and the first half of round 2 left to ATen: ReLU masks, 3x3/2 max pooling backward, the four
// adaptive average pools backward, PReLU forward / backward, Dropout2d, channel log-softmax backward, pitched copies (the pyramid
// concat is written slice by slice).  Activations are NHWC fp32 (torch channels_last storage), C % 4 == 0; every kernel is a gather
// (no atomics on the data path), so results are deterministic.
#include "df_common.cuh"
#include <stdint.h>
#include "../../include/densefusion_b200.h"

namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// out = d * (act > 0), all three with the same pixel pitch ld (cols channels per pixel)
__global__ void __launch_bounds__(256)
relu_mask_kernel2(const float* __restrict__ d, const float* __restrict__ act, float* __restrict__ out, int ld, int c4, long long rows)
{
    const long long total = rows * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c4;
        const size_t o = (size_t)r * ld + (size_t)(i - r * c4) * 4;
        const float4 g = ld4(d + o), a = ld4(act + o);
        *reinterpret_cast<float4*>(out + o) = make_float4(a.x > 0.f ? g.x : 0.f, a.y > 0.f ? g.y : 0.f, a.z > 0.f ? g.z : 0.f, a.w > 0.f ? g.w : 0.f);
    }
}

// Backward of the 3x3 / stride 2 / pad 1 max pooling: input pixel (y, x) collects dy of every window whose FIRST maximum (scan order
// dy, dx with a strict >, as ATen's max_pool2d_with_indices) is this pixel.  The window maxima are recomputed from x (no index map).
__global__ void __launch_bounds__(256)
maxpool_backward_kernel(const float* __restrict__ in, const float* __restrict__ gout, float* __restrict__ gin, int B, int H, int W, int C,
                        int Ho, int Wo)
{
    const int c4 = C >> 2;
    const long long total = (long long)B * H * W * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cq = (int)(i % c4);
        long long r = i / c4;
        const int x = (int)(r % W); r /= W;
        const int y = (int)(r % H), b = (int)(r / H);
        const float* base = in + (size_t)b * H * W * C + cq * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        // windows (yo, xo) with yo*2-1 <= y <= yo*2+1
        for (int yo = y >> 1; yo <= (y + 1) >> 1; ++yo) {
            if (yo >= Ho) continue;
            for (int xo = x >> 1; xo <= (x + 1) >> 1; ++xo) {
                if (xo >= Wo) continue;
                float m[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
                int arg[4] = {-1, -1, -1, -1};
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const int yy = yo * 2 - 1 + dy;
                    if (yy < 0 || yy >= H) continue;
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const int xx = xo * 2 - 1 + dx;
                        if (xx < 0 || xx >= W) continue;
                        const float4 v = ld4(base + ((size_t)yy * W + xx) * C);
                        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (e[k] > m[k] || e[k] != e[k]) { m[k] = e[k]; arg[k] = yy * W + xx; }
                    }
                }
                const float4 g = ld4(gout + (((size_t)b * Ho + yo) * Wo + xo) * C + cq * 4);
                const int me = y * W + x;
                if (arg[0] == me) acc.x += g.x;
                if (arg[1] == me) acc.y += g.y;
                if (arg[2] == me) acc.z += g.z;
                if (arg[3] == me) acc.w += g.w;
            }
        }
        reinterpret_cast<float4*>(gin)[i] = acc;
    }
}

// Backward of the four adaptive average pools (1, 2, 3, 6) in one pass: dx[b,y,x,:] = sum over the stages and over the bins that
// contain (y, x) -- adaptive bins overlap when the extent is not a multiple of S -- of dpool[bin] / area(bin).
// dpool (50 B, C) stage-major like df_enc_pyramid_pool's output.
__global__ void __launch_bounds__(256)
pyramid_pool_backward_kernel(const float* __restrict__ dpool, float* __restrict__ dx, int ldo, int B, int H, int W, int C)
{
    const int c4 = C >> 2;
    const long long total = (long long)B * H * W * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cq = (int)(i % c4);
        long long r = i / c4;
        const int x = (int)(r % W); r /= W;
        const int y = (int)(r % H), b = (int)(r / H);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int base = 0;                                                     // first row of the stage block, in units of B
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int S = k == 0 ? 1 : (k == 1 ? 2 : (k == 2 ? 3 : 6));
            for (int sy = 0; sy < S; ++sy) {
                const int y0 = (sy * H) / S, y1 = ((sy + 1) * H + S - 1) / S;
                if (y < y0 || y >= y1) continue;
                for (int sx = 0; sx < S; ++sx) {
                    const int x0 = (sx * W) / S, x1 = ((sx + 1) * W + S - 1) / S;
                    if (x < x0 || x >= x1) continue;
                    const float inv = 1.0f / (float)((y1 - y0) * (x1 - x0));
                    const float4 g = ld4(dpool + ((size_t)base * B + (size_t)b * S * S + sy * S + sx) * C + cq * 4);
                    acc.x += g.x * inv; acc.y += g.y * inv; acc.z += g.z * inv; acc.w += g.w * inv;
                }
            }
            base += S * S;
        }
        *reinterpret_cast<float4*>(dx + (((size_t)b * H + y) * W + x) * ldo + cq * 4) = acc;
    }
}

// log-softmax over 32 channels, backward: dx = dy - exp(y) * sum_c dy  (y = the forward's output)
__global__ void __launch_bounds__(256)
log_softmax32_backward_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, long long rows)
{
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
        float4 g[8];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { g[j] = ld4(dy + r * 32 + j * 4); s += (g[j].x + g[j].y) + (g[j].z + g[j].w); }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 v = ld4(y + r * 32 + j * 4);
            *reinterpret_cast<float4*>(dx + r * 32 + j * 4) =
                make_float4(g[j].x - __expf(v.x) * s, g[j].y - __expf(v.y) * s, g[j].z - __expf(v.z) * s, g[j].w - __expf(v.w) * s);
        }
    }
}

// PReLU with one slope (nn.PReLU()): y = x > 0 ? x : a x
__global__ void __launch_bounds__(256)
prelu_kernel(const float* __restrict__ x, const float* __restrict__ slope, float* __restrict__ y, long long n4)
{
    const float a = __ldg(slope);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = ld4(x + i * 4);
        reinterpret_cast<float4*>(y)[i] = make_float4(v.x > 0.f ? v.x : a * v.x, v.y > 0.f ? v.y : a * v.y, v.z > 0.f ? v.z : a * v.z,
                                                      v.w > 0.f ? v.w : a * v.w);
    }
}

// ... backward: dx = dy (x > 0 ? 1 : a); partial[block] = sum over the block's elements of dy x [x <= 0] (fixed grid, fixed order:
// deterministic; prelu_slope_reduce_kernel adds the partials in index order)
constexpr int PRELU_BLOCKS = 1184;
__global__ void __launch_bounds__(256)
prelu_backward_kernel(const float* __restrict__ x, const float* __restrict__ slope, const float* __restrict__ dy, float* __restrict__ dx,
                      float* __restrict__ partial, long long n4)
{
    __shared__ float red[256];
    const float a = __ldg(slope);
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = ld4(x + i * 4), g = ld4(dy + i * 4);
        reinterpret_cast<float4*>(dx)[i] = make_float4(v.x > 0.f ? g.x : a * g.x, v.y > 0.f ? g.y : a * g.y, v.z > 0.f ? g.z : a * g.z,
                                                       v.w > 0.f ? g.w : a * g.w);
        s += (v.x > 0.f ? 0.f : g.x * v.x) + (v.y > 0.f ? 0.f : g.y * v.y) + (v.z > 0.f ? 0.f : g.z * v.z) + (v.w > 0.f ? 0.f : g.w * v.w);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off >= 1; off >>= 1) {
        if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void prelu_slope_reduce_kernel(const float* __restrict__ partial, int n, float* __restrict__ out)
{
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += (double)partial[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off >= 1; off >>= 1) {
        if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)red[0];
}

// Dropout2d: one keep / drop decision per (sample, channel).  mask[i] = u_i >= p ? 1 / (1 - p) : 0 with u_i from a counter-based hash
// of (seed, call counter, i); the counter lives on the device and is advanced by the launch itself, so a captured CUDA graph draws a
// fresh mask at every replay.  state = {seed, counter} (two uint64).
__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__global__ void dropout_mask_kernel(float* __restrict__ mask, int n, float p, const unsigned long long* __restrict__ state)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t h = mix64(mix64(state[0] ^ mix64(state[1])) + (uint64_t)i);
    const float u = (float)(h >> 40) * (1.0f / 16777216.0f);
    mask[i] = u >= p ? 1.0f / (1.0f - p) : 0.0f;
}
__global__ void dropout_bump_kernel(unsigned long long* state) { state[1] += 1ull; }

// y[b, pixel, c] = x[b, pixel, c] * mask[b, c]   (forward and backward of Dropout2d)
__global__ void __launch_bounds__(256)
scale_bc_kernel(const float* __restrict__ x, const float* __restrict__ mask, float* __restrict__ y, int B, long long HW, int C)
{
    const int c4 = C >> 2;
    const long long total = (long long)B * HW * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cq = (int)(i % c4);
        const int b = (int)(i / (HW * c4));
        const float4 v = ld4(x + i * 4), m = ld4(mask + (size_t)b * C + cq * 4);
        reinterpret_cast<float4*>(y)[i] = make_float4(v.x * m.x, v.y * m.y, v.z * m.z, v.w * m.w);
    }
}

// dst[r, 0:cols] = src[r, 0:cols] with row pitches (a channel slice of an NHWC buffer)
__global__ void __launch_bounds__(256)
copy2d_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, long long rows, int c4)
{
    const long long total = rows * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c4;
        const int cq = (int)(i - r * c4);
        *reinterpret_cast<float4*>(dst + (size_t)r * ldd + cq * 4) = ld4(src + (size_t)r * lds + cq * 4);
    }
}

// out = a + b (dense)
__global__ void __launch_bounds__(256)
add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n4)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 u = ld4(a + i * 4), v = ld4(b + i * 4);
        reinterpret_cast<float4*>(out)[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
    }
}

// grid-stride launches: enough blocks for the work, at most 8 per SM of a B200 (148 SMs)
inline unsigned grid_of(long long work)
{
    const long long blocks = (work + 255) / 256;
    return (unsigned)(blocks < 1 ? 1 : (blocks > 148LL * 8 ? 148LL * 8 : blocks));
}
inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

extern "C" int df_ew_relu_mask(const float* d, const float* act, float* out, int ld, int cols, long long rows, void* stream)
{
    if (!d || !act || !out || rows <= 0 || cols <= 0 || cols % 4 || ld % 4 || ld < cols || !al16(d) || !al16(act) || !al16(out)) return DF_ERR_ARG;
    relu_mask_kernel2<<<grid_of(rows * (cols >> 2)), 256, 0, (cudaStream_t)stream>>>(d, act, out, ld, cols >> 2, rows);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_ew_maxpool_backward(const float* x, const float* gout, float* gin, int B, int H, int W, int C, void* stream)
{
    if (!x || !gout || !gin || B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 4 || !al16(x) || !al16(gout) || !al16(gin)) return DF_ERR_ARG;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    maxpool_backward_kernel<<<grid_of((long long)B * H * W * (C >> 2)), 256, 0, (cudaStream_t)stream>>>(x, gout, gin, B, H, W, C, Ho, Wo);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_ew_pyramid_pool_backward(const float* dpool, float* dx, int ldo, int B, int H, int W, int C, void* stream)
{
    if (!dpool || !dx || B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 4 || ldo % 4 || ldo < C || !al16(dpool) || !al16(dx)) return DF_ERR_ARG;
    pyramid_pool_backward_kernel<<<grid_of((long long)B * H * W * (C >> 2)), 256, 0, (cudaStream_t)stream>>>(dpool, dx, ldo, B, H, W, C);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_ew_log_softmax32_backward(const float* y, const float* dy, float* dx, long long pixels, void* stream)
{
    if (!y || !dy || !dx || pixels <= 0 || !al16(y) || !al16(dy) || !al16(dx)) return DF_ERR_ARG;
    log_softmax32_backward_kernel<<<grid_of(pixels), 256, 0, (cudaStream_t)stream>>>(y, dy, dx, pixels);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_ew_prelu(const float* x, const float* slope, float* y, long long n, void* stream)
{
    if (!x || !slope || !y || n <= 0 || n % 4 || !al16(x) || !al16(y)) return DF_ERR_ARG;
    prelu_kernel<<<grid_of(n >> 2), 256, 0, (cudaStream_t)stream>>>(x, slope, y, n >> 2);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_ew_prelu_scratch_floats(void) { return PRELU_BLOCKS; }

extern "C" int df_ew_prelu_backward(const float* x, const float* slope, const float* dy, float* dx, float* dslope, float* scratch,
                                    long long n, void* stream)
{
    if (!x || !slope || !dy || !dx || !dslope || !scratch || n <= 0 || n % 4 || !al16(x) || !al16(dy) || !al16(dx)) return DF_ERR_ARG;
    prelu_backward_kernel<<<PRELU_BLOCKS, 256, 0, (cudaStream_t)stream>>>(x, slope, dy, dx, scratch, n >> 2);
    prelu_slope_reduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(scratch, PRELU_BLOCKS, dslope);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_ew_dropout_mask(float* mask, int n, float p, unsigned long long* state, void* stream)
{
    if (!mask || !state || n <= 0 || !(p >= 0.0f && p < 1.0f)) return DF_ERR_ARG;
    dropout_mask_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(mask, n, p, state);
    dropout_bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_ew_scale_bc(const float* x, const float* mask, float* y, int B, long long HW, int C, void* stream)
{
    if (!x || !mask || !y || B <= 0 || HW <= 0 || C <= 0 || C % 4 || !al16(x) || !al16(mask) || !al16(y)) return DF_ERR_ARG;
    scale_bc_kernel<<<grid_of((long long)B * HW * (C >> 2)), 256, 0, (cudaStream_t)stream>>>(x, mask, y, B, HW, C);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_ew_copy2d(const float* src, int lds, float* dst, int ldd, long long rows, int cols, void* stream)
{
    if (!src || !dst || rows <= 0 || cols <= 0 || cols % 4 || lds % 4 || ldd % 4 || lds < cols || ldd < cols || !al16(src) || !al16(dst))
        return DF_ERR_ARG;
    copy2d_kernel<<<grid_of(rows * (cols >> 2)), 256, 0, (cudaStream_t)stream>>>(src, lds, dst, ldd, rows, cols >> 2);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_ew_add(const float* a, const float* b, float* out, long long n, void* stream)
{
    if (!a || !b || !out || n <= 0 || n % 4 || !al16(a) || !al16(b) || !al16(out)) return DF_ERR_ARG;
    add_kernel<<<grid_of(n >> 2), 256, 0, (cudaStream_t)stream>>>(a, b, out, n >> 2);
    DF_RETURN_LAST_ERROR();
}
