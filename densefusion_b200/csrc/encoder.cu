// NHWC helper kernels of the colour encoder (reference: lib/extractors.py:78-124, lib/pspnet.py:7-77).  The
// convolutions themselves are implicit GEMMs on the tcgen05 kernel (df_conv_tc, gemm_tc.cu); what is left is
// HBM-bound data movement: im2col for the three stride-2 layers, max / average pooling, bilinear resizing into a
// channel slice of a wider buffer (the pyramid concat is never materialised separately) and the channel log-softmax.
// All activations are channels-last so that a pixel's channels are one contiguous, float4-readable run.
#include "df_common.cuh"
#include <stdlib.h>
#include "../../include/densefusion_b200.h"

namespace {

// conv1 (7x7, stride 2, pad 3) as a GEMM: A[pixel, c*49 + ky*7 + kx] from the NCHW image, zero-padded to ldk columns.
// One thread per (pixel, 4 consecutive k): one float4 store.  The (c, ky, kx) split of k comes from a table every block
// builds in shared memory (c | ky << 8 | kx << 16, 0xffffffff for the zero padding): the lanes of a warp index it with 32
// different k, which a __constant__ table would serialise (measured 470 -> 130 us for 64 crops of 160x160) -- and a constant
// table needs a host upload whose DMA from pageable memory may still be in flight when the first kernel on a non-blocking
// stream reads it.
__global__ void __launch_bounds__(256)
im2col_conv1_kernel(const float* __restrict__ img, float* __restrict__ A, int B, int H, int W, int Ho, int Wo, int ldk)
{
    __shared__ uint32_t s_k[192];
    if (threadIdx.x < 192) {
        const int k = threadIdx.x, c = k / 49, r = k - c * 49;
        s_k[k] = k < 147 ? (uint32_t)(c | ((r / 7) << 8) | ((r % 7) << 16)) : 0xffffffffu;
    }
    __syncthreads();
    const unsigned kq_per_pix = (unsigned)ldk >> 2;
    const unsigned total = (unsigned)B * Ho * Wo * kq_per_pix;          // < 2^31 (checked by the launcher)
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned kq = i % kq_per_pix;
        const unsigned pix = i / kq_per_pix;
        const int xo = (int)(pix % (unsigned)Wo);
        const unsigned r = pix / (unsigned)Wo;
        const int yo = (int)(r % (unsigned)Ho), b = (int)(r / (unsigned)Ho);
        const float* base = img + (size_t)b * 3 * H * W;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned k = kq * 4 + e;
            const uint32_t t = k < 192 ? s_k[k] : 0xffffffffu;
            float x = 0.0f;
            if (t != 0xffffffffu) {
                const int c = t & 0xff, ky = (t >> 8) & 0xff, kx = (t >> 16) & 0xff;
                const int yy = yo * 2 - 3 + ky, xx = xo * 2 - 3 + kx;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) x = __ldg(base + ((size_t)c * H + yy) * W + xx);
            }
            v[e] = x;
        }
        reinterpret_cast<float4*>(A)[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// 3x3 / stride 2 / pad 1 max pooling, NHWC, float4 over channels
// (IDX: unsigned when the element count fits 31 bits -- every bench shape -- else long long.  The ncu capture of the last session put both
// this kernel and im2col_s2 at 66-75% issue-active with DRAM far from busy: the 64-bit divisions of the index decode, ~100 instructions
// each, were most of what the threads executed.)
template <typename IDX>
__global__ void __launch_bounds__(256)
maxpool_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int H, int W, int C, int Ho, int Wo)
{
    const IDX c4 = (IDX)(C >> 2);
    const IDX total = (IDX)B * (IDX)Ho * (IDX)Wo * c4;
    for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IDX)gridDim.x * blockDim.x) {
        const int cq = (int)(i % c4);
        IDX r = i / c4;
        const int xo = (int)(r % (IDX)Wo); r /= (IDX)Wo;
        const int yo = (int)(r % (IDX)Ho), b = (int)(r / (IDX)Ho);
        float4 m = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int y = yo * 2 - 1 + dy;
            if (y < 0 || y >= H) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int x = xo * 2 - 1 + dx;
                if (x < 0 || x >= W) continue;
                const float4 v = __ldg(reinterpret_cast<const float4*>(in + (((size_t)b * H + y) * W + x) * C) + cq);
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
        reinterpret_cast<float4*>(out)[i] = m;
    }
}

// 3x3 / stride 2 / pad 1 patches, NHWC -> A[pixel, tap*C + c]  (tap-major, like the repacked conv weights)
template <typename IDX>
__global__ void __launch_bounds__(256)
im2col_s2_kernel(const float* __restrict__ in, float* __restrict__ A, int B, int H, int W, int C, int Ho, int Wo)
{
    const IDX c4 = (IDX)(C >> 2);
    const IDX total = (IDX)B * (IDX)Ho * (IDX)Wo * 9 * c4;
    for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IDX)gridDim.x * blockDim.x) {
        const int cq = (int)(i % c4);
        IDX r = i / c4;
        const int tap = (int)(r % 9); r /= 9;
        const int xo = (int)(r % (IDX)Wo); r /= (IDX)Wo;
        const int yo = (int)(r % (IDX)Ho), b = (int)(r / (IDX)Ho);
        const int y = yo * 2 - 1 + tap / 3, x = xo * 2 - 1 + tap % 3;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(reinterpret_cast<const float4*>(in + (((size_t)b * H + y) * W + x) * C) + cq);
        reinterpret_cast<float4*>(A)[i] = v;
    }
}

// Transpose of im2col_s2 in gather form (deterministic, no atomics) -- the data gradient of a 3x3 / stride 2 / pad 1
// convolution from dA[pixel_out, tap*C + c]: input pixel (y, x) collects tap (ky, kx) of output pixel ((y+1-ky)/2, (x+1-kx)/2)
// wherever those are integers inside the output map.
__global__ void __launch_bounds__(256)
col2im_s2_kernel(const float* __restrict__ dA, float* __restrict__ dx, int B, int H, int W, int C, int Ho, int Wo)
{
    const int c4 = C >> 2;
    const long long total = (long long)B * H * W * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cq = (int)(i % c4);
        long long r = i / c4;
        const int x = (int)(r % W); r /= W;
        const int y = (int)(r % H), b = (int)(r / H);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int ty = y + 1 - ky;
            if (ty < 0 || (ty & 1) || (ty >> 1) >= Ho) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int tx = x + 1 - kx;
                if (tx < 0 || (tx & 1) || (tx >> 1) >= Wo) continue;
                const float4 v = __ldg(reinterpret_cast<const float4*>(dA + ((((size_t)b * Ho + (ty >> 1)) * Wo + (tx >> 1)) * 9 + ky * 3 + kx) * C) + cq);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        reinterpret_cast<float4*>(dx)[i] = acc;
    }
}

// nn.AdaptiveAvgPool2d((S,S)) on NHWC (pixel pitch ldi): bin i covers [floor(i*H/S), ceil((i+1)*H/S))
__global__ void __launch_bounds__(256)
adaptive_avgpool_kernel(const float* __restrict__ in, int ldi, float* __restrict__ out, int B, int H, int W, int C, int S)
{
    const long long total = (long long)B * S * S * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long r = i / C;
        const int sx = (int)(r % S); r /= S;
        const int sy = (int)(r % S), b = (int)(r / S);
        const int y0 = (sy * H) / S, y1 = ((sy + 1) * H + S - 1) / S;
        const int x0 = (sx * W) / S, x1 = ((sx + 1) * W + S - 1) / S;
        float s = 0.0f;
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) s += __ldg(in + (((size_t)b * H + y) * W + x) * ldi + c);
        out[i] = s / (float)((y1 - y0) * (x1 - x0));
    }
}

// bilinear resize NHWC (pixel pitches ldi / ldo, so the output may be a channel slice of a wider buffer); index and
// weight arithmetic as ATen's area_pixel_compute_source_index, same as upsample.cu
__global__ void __launch_bounds__(256)
upsample_nhwc_kernel(const float* __restrict__ in, int ldi, float* __restrict__ out, int ldo, int B, int hin, int win,
                     int hout, int wout, int C, float rh, float rw, int align)
{
    const unsigned c4 = C >> 2;
    const unsigned total = (unsigned)B * hout * wout * c4;              // < 2^31 (checked by the launcher; no wrap in the stride loop)
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int cq = (int)(i % c4);
        unsigned r = i / c4;
        const int x = (int)(r % (unsigned)wout); r /= (unsigned)wout;
        const int y = (int)(r % (unsigned)hout), b = (int)(r / (unsigned)hout);
        float sy, sx;
        if (align) { sy = rh * y; sx = rw * x; }
        else {
            sy = rh * (y + 0.5f) - 0.5f; sy = sy < 0.f ? 0.f : sy;
            sx = rw * (x + 0.5f) - 0.5f; sx = sx < 0.f ? 0.f : sx;
        }
        const int y0 = (int)sy, x0 = (int)sx;
        const int yp = y0 < hin - 1 ? 1 : 0, xp = x0 < win - 1 ? 1 : 0;
        const float ly1 = sy - y0, ly0 = 1.0f - ly1, lx1 = sx - x0, lx0 = 1.0f - lx1;
        const float* p = in + (((size_t)b * hin + y0) * win + x0) * ldi + cq * 4;
        const float4 v00 = __ldg(reinterpret_cast<const float4*>(p));
        const float4 v01 = __ldg(reinterpret_cast<const float4*>(p + (size_t)xp * ldi));
        const float4 v10 = __ldg(reinterpret_cast<const float4*>(p + (size_t)yp * win * ldi));
        const float4 v11 = __ldg(reinterpret_cast<const float4*>(p + ((size_t)yp * win + xp) * ldi));
        float4 o;
        o.x = ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
        o.y = ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
        o.z = ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
        o.w = ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
        *reinterpret_cast<float4*>(out + (((size_t)b * hout + y) * wout + x) * ldo + cq * 4) = o;
    }
}

// ---- folded pyramid (PSP) ----
// The pyramid branch of lib/pspnet.py:17-24 is  relu(Wb . cat[up(S_s . pool_s(f)) for s in 1,2,3,6; f] + b).  The 1x1 convolutions
// are pointwise and the bilinear resize is linear, so  Wb_s . up(S_s . pool_s(f)) = up((Wb_s S_s) . pool_s(f)) : the four
// products run at the pooled resolution (50 cells per crop) and only the 512 `f` channels go through the full-resolution
// GEMM.  pyramid_pool: the four adaptive average pools in one pass -> (50 B, C) stage-major: rows [B x 1 | B x 4 | B x 9 | B x 36],
// so that every stage is one dense GEMM operand.
__global__ void __launch_bounds__(256)
pyramid_pool_kernel(const float* __restrict__ in, int ldi, float* __restrict__ out, int B, int H, int W, int C, unsigned first_row)
{
    const unsigned c4 = C >> 2;
    const unsigned total = ((unsigned)B * 50u - first_row) * c4;
    for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += gridDim.x * blockDim.x) {
        const unsigned i = i0 + first_row * c4;
        const int cq = (int)(i % c4);
        const unsigned r = i / c4;                                            // stage-major row: [B x 1 | B x 4 | B x 9 | B x 36]
        int S, q;
        if (r < (unsigned)B) { S = 1; q = (int)r; } else if (r < 5u * B) { S = 2; q = (int)r - B; }
        else if (r < 14u * B) { S = 3; q = (int)r - 5 * B; } else { S = 6; q = (int)r - 14 * B; }
        const int b = q / (S * S), k = q - b * S * S, sy = k / S, sx = k - sy * S;
        const int y0 = (sy * H) / S, y1 = ((sy + 1) * H + S - 1) / S;          // nn.AdaptiveAvgPool2d bins
        const int x0 = (sx * W) / S, x1 = ((sx + 1) * W + S - 1) / S;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(in + (((size_t)b * H + y) * W + x) * ldi) + cq);
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
        const float n = (float)((y1 - y0) * (x1 - x0));
        reinterpret_cast<float4*>(out)[i] = make_float4(a.x / n, a.y / n, a.z / n, a.w / n);
    }
}

// The same pools with the pixels of a cell shared out: one CTA per (stage cell, group of 32 channel quads), its 8 warps take the cell's
// pixels round robin (row-major inside the bin) and the partial sums meet in shared memory, added in warp order -- a fixed order, so
// the result is deterministic.  The one-thread-per-cell kernel above walks the 1 x 1 stage's whole map (400 pixels at 20 x 20) in
// one thread: a latency chain that sets the launch time (0.04 ms for 52 MB that sit in L2).
__global__ void __launch_bounds__(256)
pyramid_pool_split_kernel(const float* __restrict__ in, int ldi, float* __restrict__ out, int B, int H, int W, int C, unsigned rows)
{
    __shared__ float4 s_part[8][32];
    const int c4 = C >> 2, groups = (c4 + 31) >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned total = rows * (unsigned)groups;
    for (unsigned u = blockIdx.x; u < total; u += gridDim.x) {
        const unsigned r = u / (unsigned)groups;                              // stage-major row: [B x 1 | B x 4 | B x 9 | B x 36]
        const int cq = (int)(u - r * (unsigned)groups) * 32 + lane;
        int S, q;
        if (r < (unsigned)B) { S = 1; q = (int)r; } else if (r < 5u * B) { S = 2; q = (int)r - B; }
        else if (r < 14u * B) { S = 3; q = (int)r - 5 * B; } else { S = 6; q = (int)r - 14 * B; }
        const int b = q / (S * S), k = q - b * S * S, sy = k / S, sx = k - sy * S;
        const int y0 = (sy * H) / S, y1 = ((sy + 1) * H + S - 1) / S;          // nn.AdaptiveAvgPool2d bins
        const int x0 = (sx * W) / S, x1 = ((sx + 1) * W + S - 1) / S;
        const int bw = x1 - x0, npix = (y1 - y0) * bw;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cq < c4) {
#pragma unroll 4
            for (int j = warp; j < npix; j += 8) {
                const int jy = j / bw, jx = j - jy * bw;
                const float4 v = __ldg(reinterpret_cast<const float4*>(in + (((size_t)b * H + y0 + jy) * W + x0 + jx) * ldi) + cq);
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
        }
        s_part[warp][lane] = a;
        __syncthreads();
        if (warp == 0 && cq < c4) {
            float4 t = s_part[0][lane];
#pragma unroll
            for (int p = 1; p < 8; ++p) { const float4 v = s_part[p][lane]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
            const float n = (float)npix;
            reinterpret_cast<float4*>(out)[(size_t)r * c4 + cq] = make_float4(t.x / n, t.y / n, t.z / n, t.w / n);
        }
        __syncthreads();
    }
}

// pyramid_sum: out[b,y,x,:] = Y[b,0,:] + sum over s in (2,3,6) of the bilinear (align_corners = False) resize of the s x s
// cells of Y (50 B, C; stage-major like pyramid_pool's output) to H x W -- the residual operand of the bottleneck GEMM.
// Interpolation arithmetic as upsample_nhwc_kernel.
__global__ void __launch_bounds__(256)
pyramid_sum_kernel(const float* __restrict__ Y, float* __restrict__ out, int ldo, int B, int H, int W, int C)
{
    const unsigned c4 = C >> 2;
    const unsigned total = (unsigned)B * H * W * c4;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int cq = (int)(i % c4);
        unsigned r = i / c4;
        const int x = (int)(r % (unsigned)W); r /= (unsigned)W;
        const int y = (int)(r % (unsigned)H), b = (int)(r / (unsigned)H);
        float4 acc = __ldg(reinterpret_cast<const float4*>(Y + (size_t)b * C) + cq);
        int base = 1;                                                          // first row of the stage block, in units of B
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int S = k == 0 ? 2 : (k == 1 ? 3 : 6);
            float sy = ((float)S / H) * (y + 0.5f) - 0.5f; sy = sy < 0.f ? 0.f : sy;
            float sx = ((float)S / W) * (x + 0.5f) - 0.5f; sx = sx < 0.f ? 0.f : sx;
            const int y0 = (int)sy, x0 = (int)sx;
            const int yp = y0 < S - 1 ? 1 : 0, xp = x0 < S - 1 ? 1 : 0;
            const float ly1 = sy - y0, ly0 = 1.0f - ly1, lx1 = sx - x0, lx0 = 1.0f - lx1;
            const float* p = Y + ((size_t)base * B + (size_t)b * S * S + y0 * S + x0) * C + cq * 4;
            const float4 v00 = __ldg(reinterpret_cast<const float4*>(p));
            const float4 v01 = __ldg(reinterpret_cast<const float4*>(p + (size_t)xp * C));
            const float4 v10 = __ldg(reinterpret_cast<const float4*>(p + (size_t)yp * S * C));
            const float4 v11 = __ldg(reinterpret_cast<const float4*>(p + ((size_t)yp * S + xp) * C));
            acc.x += ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
            acc.y += ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
            acc.z += ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
            acc.w += ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
            base += S * S;
        }
        *reinterpret_cast<float4*>(out + (((size_t)b * H + y) * W + x) * ldo + cq * 4) = acc;
    }
}

// pyramid_sum with one thread per (crop, row, channel quad) walking the row: the four corner vectors of a stage change only when
// the source cell does (2 + 3 + 6 times along a row), so they stay in registers in between -- ~2 instead of 12 loads per output
// float4.  The per-pixel kernel above re-reads its 12 vectors for every output from L2 (a CTA is one pixel: nothing for L1 to reuse),
// 12 bytes through the SM's port per byte written.  Same arithmetic per element, expression for expression.
__global__ void __launch_bounds__(256)
pyramid_sum_rows_kernel(const float* __restrict__ Y, float* __restrict__ out, int ldo, int B, int H, int W, int C)
{
    const unsigned c4 = C >> 2;
    const unsigned total = (unsigned)B * H * c4;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int cq = (int)(i % c4);
        const unsigned r = i / c4;
        const int y = (int)(r % (unsigned)H), b = (int)(r / (unsigned)H);
        const float4 g = __ldg(reinterpret_cast<const float4*>(Y + (size_t)b * C) + cq);
        const float* rowp[3];                                                  // row y0 of the stage's cells of this crop
        size_t rstep[3];                                                       // to row y0 + 1 (0 on the last row)
        float ly0[3], ly1[3];
        int base = 1;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int S = k == 0 ? 2 : (k == 1 ? 3 : 6);
            float sy = ((float)S / H) * (y + 0.5f) - 0.5f; sy = sy < 0.f ? 0.f : sy;
            const int y0 = (int)sy;
            const int yp = y0 < S - 1 ? 1 : 0;
            ly1[k] = sy - y0; ly0[k] = 1.0f - ly1[k];
            rowp[k] = Y + ((size_t)base * B + (size_t)b * S * S + y0 * S) * C + cq * 4;
            rstep[k] = (size_t)yp * S * C;
            base += S * S;
        }
        int cx0[3] = {-1, -1, -1};
        float4 v00[3], v01[3], v10[3], v11[3];
        float* orow = out + (((size_t)b * H + y) * W) * ldo + cq * 4;
        for (int x = 0; x < W; ++x) {
            float4 acc = g;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int S = k == 0 ? 2 : (k == 1 ? 3 : 6);
                float sx = ((float)S / W) * (x + 0.5f) - 0.5f; sx = sx < 0.f ? 0.f : sx;
                const int x0 = (int)sx;
                const float lx1 = sx - x0, lx0 = 1.0f - lx1;
                if (x0 != cx0[k]) {
                    const int xp = x0 < S - 1 ? 1 : 0;
                    const float* p = rowp[k] + (size_t)x0 * C;
                    v00[k] = __ldg(reinterpret_cast<const float4*>(p));
                    v01[k] = __ldg(reinterpret_cast<const float4*>(p + (size_t)xp * C));
                    v10[k] = __ldg(reinterpret_cast<const float4*>(p + rstep[k]));
                    v11[k] = __ldg(reinterpret_cast<const float4*>(p + rstep[k] + (size_t)xp * C));
                    cx0[k] = x0;
                }
                acc.x += ly0[k] * (lx0 * v00[k].x + lx1 * v01[k].x) + ly1[k] * (lx0 * v10[k].x + lx1 * v11[k].x);
                acc.y += ly0[k] * (lx0 * v00[k].y + lx1 * v01[k].y) + ly1[k] * (lx0 * v10[k].y + lx1 * v11[k].y);
                acc.z += ly0[k] * (lx0 * v00[k].z + lx1 * v01[k].z) + ly1[k] * (lx0 * v10[k].z + lx1 * v11[k].z);
                acc.w += ly0[k] * (lx0 * v00[k].w + lx1 * v01[k].w) + ly1[k] * (lx0 * v10[k].w + lx1 * v11[k].w);
            }
            *reinterpret_cast<float4*>(orow + (size_t)x * ldo) = acc;
        }
    }
}

// pyramid_sum, third form (default): the row walk of pyramid_sum_rows_kernel with the crop's 50 cell vectors staged in shared memory
// first -- one CTA per (crop, 32 channel quads), 25.6 KB, five warps taking the rows round robin (H = 10 / 15 / 20: whole rounds).  A
// change of source cell then costs a shared-memory read instead of an L2 round trip in the middle of a serial walk (rows kernel: ~11 such
// stalls per row at 16 warps per SM, 0.166 ms per step against 0.209 for the per-pixel kernel; this one 0.113,
// profiles/r2_s5_small_kernels.txt).  Same arithmetic per element.
constexpr int PS_WARPS = 5;
__global__ void __launch_bounds__(PS_WARPS * 32)
pyramid_sum_smem_kernel(const float* __restrict__ Y, float* __restrict__ out, int ldo, int B, int H, int W, int C)
{
    __shared__ float4 s_y[50 * 32];
    const int c4 = C >> 2, slices = (c4 + 31) >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned total = (unsigned)B * (unsigned)slices;
    for (unsigned u = blockIdx.x; u < total; u += gridDim.x) {
        const int b = (int)(u / (unsigned)slices);
        const int cq = (int)(u - (unsigned)b * (unsigned)slices) * 32 + lane;
        __syncthreads();                                                       // (the previous crop's readers are done)
        for (int j = warp; j < 50; j += PS_WARPS) {                            // stage-major rows of Y: [B x 1 | B x 4 | B x 9 | B x 36]
            const size_t row = j == 0 ? (size_t)b : j < 5 ? (size_t)B + (size_t)b * 4 + (j - 1)
                               : j < 14 ? (size_t)5 * B + (size_t)b * 9 + (j - 5) : (size_t)14 * B + (size_t)b * 36 + (j - 14);
            if (cq < c4) s_y[j * 32 + lane] = __ldg(reinterpret_cast<const float4*>(Y + row * C) + cq);
        }
        __syncthreads();
        if (cq >= c4) continue;
        const float4 g = s_y[lane];
        for (int y = warp; y < H; y += PS_WARPS) {
            int rowc[3], rstep[3];                                             // first cell of row y0 of the stage / step to row y0 + 1
            float ly0[3], ly1[3];
            int base = 1;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int S = k == 0 ? 2 : (k == 1 ? 3 : 6);
                float sy = ((float)S / H) * (y + 0.5f) - 0.5f; sy = sy < 0.f ? 0.f : sy;
                const int y0 = (int)sy;
                const int yp = y0 < S - 1 ? 1 : 0;
                ly1[k] = sy - y0; ly0[k] = 1.0f - ly1[k];
                rowc[k] = base + y0 * S;
                rstep[k] = yp * S;
                base += S * S;
            }
            int cx0[3] = {-1, -1, -1};
            float4 v00[3], v01[3], v10[3], v11[3];
            float* orow = out + (((size_t)b * H + y) * W) * ldo + cq * 4;
            for (int x = 0; x < W; ++x) {
                float4 acc = g;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int S = k == 0 ? 2 : (k == 1 ? 3 : 6);
                    float sx = ((float)S / W) * (x + 0.5f) - 0.5f; sx = sx < 0.f ? 0.f : sx;
                    const int x0 = (int)sx;
                    const float lx1 = sx - x0, lx0 = 1.0f - lx1;
                    if (x0 != cx0[k]) {
                        const int xp = x0 < S - 1 ? 1 : 0;
                        const float4* p = s_y + (rowc[k] + x0) * 32 + lane;
                        v00[k] = p[0];
                        v01[k] = p[xp * 32];
                        v10[k] = p[rstep[k] * 32];
                        v11[k] = p[(rstep[k] + xp) * 32];
                        cx0[k] = x0;
                    }
                    acc.x += ly0[k] * (lx0 * v00[k].x + lx1 * v01[k].x) + ly1[k] * (lx0 * v10[k].x + lx1 * v11[k].x);
                    acc.y += ly0[k] * (lx0 * v00[k].y + lx1 * v01[k].y) + ly1[k] * (lx0 * v10[k].y + lx1 * v11[k].y);
                    acc.z += ly0[k] * (lx0 * v00[k].z + lx1 * v01[k].z) + ly1[k] * (lx0 * v10[k].z + lx1 * v11[k].z);
                    acc.w += ly0[k] * (lx0 * v00[k].w + lx1 * v01[k].w) + ly1[k] * (lx0 * v10[k].w + lx1 * v11[k].w);
                }
                *reinterpret_cast<float4*>(orow + (size_t)x * ldo) = acc;
            }
        }
    }
}

// ---- 3x3 convolution after a x2 bilinear resize, evaluated at the LOW resolution (decoder stages, lib/pspnet.py:27-37) ----
// conv3x3(up(x)) = sum_tap shift_tap(W_tap . up(x)) = sum_tap shift_tap(up(W_tap . x)): a 1x1 convolution commutes with the
// resize, so Z = x . [W_0 .. W_8] is ONE GEMM over the low-resolution pixels (a quarter of the full-resolution rows: 4x fewer
// FLOPs than convolving the resized map) and this kernel finishes the layer: every output pixel sums, over the nine taps, the
// bilinear (align_corners) sample of Z's tap slice at the tap's shifted position -- positions outside the resized map are the
// convolution's zero padding -- then adds the bias and applies PReLU.  Z (B,h,w,9*C) tap-major; out (B,2h,2w,C).
// Interpolation arithmetic as upsample_nhwc_kernel.
__global__ void __launch_bounds__(256)
upconv_finish_kernel(const float* __restrict__ Z, int ldz, const float* __restrict__ bias, const float* __restrict__ prelu,
                     float* __restrict__ out, int ldo, int B, int h, int w, int C, float rh, float rw)
{
    const int H = 2 * h, W = 2 * w;
    const unsigned c4 = C >> 2;
    const unsigned total = (unsigned)B * H * W * c4;
    const float slope = __ldg(prelu);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int cq = (int)(i % c4);
        unsigned r = i / c4;
        const int X = (int)(r % (unsigned)W); r /= (unsigned)W;
        const int Y = (int)(r % (unsigned)H), b = (int)(r / (unsigned)H);
        float4 acc = bias ? __ldg(reinterpret_cast<const float4*>(bias) + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float* zb = Z + (size_t)b * h * w * ldz + cq * 4;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = Y + ky - 1;
            if (yy < 0 || yy >= H) continue;
            const float sy = rh * yy;
            const int y0 = (int)sy, yp = y0 < h - 1 ? 1 : 0;
            const float ly1 = sy - y0, ly0 = 1.0f - ly1;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = X + kx - 1;
                if (xx < 0 || xx >= W) continue;
                const float sx = rw * xx;
                const int x0 = (int)sx, xp = x0 < w - 1 ? 1 : 0;
                const float lx1 = sx - x0, lx0 = 1.0f - lx1;
                const float* p = zb + ((size_t)y0 * w + x0) * ldz + (ky * 3 + kx) * C;
                const float4 v00 = __ldg(reinterpret_cast<const float4*>(p));
                const float4 v01 = __ldg(reinterpret_cast<const float4*>(p + (size_t)xp * ldz));
                const float4 v10 = __ldg(reinterpret_cast<const float4*>(p + (size_t)yp * w * ldz));
                const float4 v11 = __ldg(reinterpret_cast<const float4*>(p + ((size_t)yp * w + xp) * ldz));
                acc.x += ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
                acc.y += ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
                acc.z += ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
                acc.w += ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
            }
        }
        acc.x = acc.x > 0.f ? acc.x : slope * acc.x; acc.y = acc.y > 0.f ? acc.y : slope * acc.y;
        acc.z = acc.z > 0.f ? acc.z : slope * acc.z; acc.w = acc.w > 0.f ? acc.w : slope * acc.w;
        *reinterpret_cast<float4*>(out + (((size_t)b * H + Y) * W + X) * ldo + cq * 4) = acc;
    }
}

// The same layer with the low-resolution footprint staged in shared memory.  upconv_finish_kernel issues 36 global float4
// loads per output float4 and is bound by L1 wavefronts (ncu: 66% SM throughput, Z read from DRAM exactly once); here a CTA
// owns an 8 x 8 patch of output pixels x 32 channels, whose nine taps sample at most 7 x 7 low-resolution pixels: that
// footprint (<= 7*7*9 rows of 128 B) is loaded once, coalesced (5 instead of 36 global loads per output), and the 36 corner
// reads become conflict-free LDS.128 (the 8 lanes of a quarter warp read one 128-byte row).  Arithmetic and its order are
// those of upconv_finish_kernel, so the results are bit-identical.
constexpr int UF_T = 8;                        // output patch edge
constexpr int UF_FP = 7;                       // low-resolution footprint edge (upper bound, see above)
constexpr int UF_SMEM = UF_FP * UF_FP * 9 * 128;

// LANES = float4 lanes per pixel a CTA handles: 8 (32 channels, 56 KB of footprint: 4 CTAs per SM; production) or 4 (16 channels,
// 28 KB: 8 CTAs per SM; A/B variant)
template <int LANES>
__global__ void __launch_bounds__(256)
upconv_finish_smem_kernel(const float* __restrict__ Z, int ldz, const float* __restrict__ bias, const float* __restrict__ prelu,
                          float* __restrict__ out, int ldo, int h, int w, int C, float rh, float rw, int tiles_x)
{
    extern __shared__ __align__(16) uint8_t uf_smem[];
    float4* s = reinterpret_cast<float4*>(uf_smem);                 // [(ry * cols + rx) * 9 + tap][LANES]
    constexpr int PIX_PER_PASS = 256 / LANES;
    const int H = 2 * h, W = 2 * w;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int Y0 = ty * UF_T, X0 = tx * UF_T, c0 = blockIdx.y * (LANES * 4), b = blockIdx.z;
    // low-resolution rows / columns the patch's taps can touch (same index arithmetic as the compute phase below)
    const int ya = max(Y0 - 1, 0), yb = min(Y0 + UF_T, H - 1), xa = max(X0 - 1, 0), xb = min(X0 + UF_T, W - 1);
    const int y_lo = (int)(rh * ya), x_lo = (int)(rw * xa);
    const int y_hi = min((int)(rh * yb) + 1, h - 1), x_hi = min((int)(rw * xb) + 1, w - 1);
    const int rows = y_hi - y_lo + 1, cols = x_hi - x_lo + 1;       // <= UF_FP each
    const float* zb = Z + (size_t)b * h * w * ldz + c0;
    for (int i = threadIdx.x; i < rows * cols * 9 * LANES; i += 256) {
        const int c = i % LANES, pt = i / LANES;
        const int tap = pt % 9, px = pt / 9;
        const int ry = px / cols, rx = px - ry * cols;
        s[i] = __ldg(reinterpret_cast<const float4*>(zb + ((size_t)(y_lo + ry) * w + (x_lo + rx)) * ldz + tap * C) + c);
    }
    __syncthreads();
    const float slope = __ldg(prelu);
    const int cq = threadIdx.x % LANES;
    const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias + c0) + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int pass = 0; pass < UF_T * UF_T / PIX_PER_PASS; ++pass) {
        const int pix = pass * PIX_PER_PASS + (threadIdx.x / LANES);
        const int Y = Y0 + pix / UF_T, X = X0 + pix % UF_T;
        if (Y >= H || X >= W) continue;
        float4 acc = bv;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = Y + ky - 1;
            if (yy < 0 || yy >= H) continue;
            const float sy = rh * yy;
            const int y0 = (int)sy, yp = y0 < h - 1 ? 1 : 0;
            const float ly1 = sy - y0, ly0 = 1.0f - ly1;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = X + kx - 1;
                if (xx < 0 || xx >= W) continue;
                const float sx = rw * xx;
                const int x0 = (int)sx, xp = x0 < w - 1 ? 1 : 0;
                const float lx1 = sx - x0, lx0 = 1.0f - lx1;
                const float4* p = s + (((y0 - y_lo) * cols + (x0 - x_lo)) * 9 + ky * 3 + kx) * LANES + cq;
                const float4 v00 = p[0];
                const float4 v01 = p[xp * 9 * LANES];
                const float4 v10 = p[yp * cols * 9 * LANES];
                const float4 v11 = p[(yp * cols + xp) * 9 * LANES];
                acc.x += ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
                acc.y += ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
                acc.z += ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
                acc.w += ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
            }
        }
        acc.x = acc.x > 0.f ? acc.x : slope * acc.x; acc.y = acc.y > 0.f ? acc.y : slope * acc.y;
        acc.z = acc.z > 0.f ? acc.z : slope * acc.z; acc.w = acc.w > 0.f ? acc.w : slope * acc.w;
        *reinterpret_cast<float4*>(out + (((size_t)b * H + Y) * W + X) * ldo + c0 + cq * 4) = acc;
    }
}

// The same kernel with fewer instructions per output (the kernel is issue-bound: ncu 77% issue-active, and its SASS spends ~46
// instructions per tap, 28 of them floating point, for 4 shared-memory reads): the row / column quantities of the three ky / kx are
// computed once per pixel (index into the footprint, step to the neighbour, the two 1-D weights), the four corner weights of a tap are
// products of those (4 multiplies per tap, shared by the 4 channels of the lane) and every corner is ONE fused multiply-add per
// channel: 16 FFMA + 4 FMUL per tap instead of 12 FMUL + 12 FFMA + 4 FADD.  The rounding differs from upconv_finish_kernel's
// interpolate-then-add order in the last bit (both are within 5e-6 of the float64 layer, tests/test_encoder_gpu.py).
template <int LANES>
__global__ void __launch_bounds__(256)
upconv_finish_smem_fma_kernel(const float* __restrict__ Z, int ldz, const float* __restrict__ bias, const float* __restrict__ prelu,
                              float* __restrict__ out, int ldo, int h, int w, int C, float rh, float rw, int tiles_x)
{
    extern __shared__ __align__(16) uint8_t uf_smem[];
    float4* s = reinterpret_cast<float4*>(uf_smem);                 // [(ry * cols + rx) * 9 + tap][LANES]
    constexpr int PIX_PER_PASS = 256 / LANES;
    const int H = 2 * h, W = 2 * w;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int Y0 = ty * UF_T, X0 = tx * UF_T, c0 = blockIdx.y * (LANES * 4), b = blockIdx.z;
    const int ya = max(Y0 - 1, 0), yb = min(Y0 + UF_T, H - 1), xa = max(X0 - 1, 0), xb = min(X0 + UF_T, W - 1);
    const int y_lo = (int)(rh * ya), x_lo = (int)(rw * xa);
    const int y_hi = min((int)(rh * yb) + 1, h - 1), x_hi = min((int)(rw * xb) + 1, w - 1);
    const int rows = y_hi - y_lo + 1, cols = x_hi - x_lo + 1;       // <= UF_FP each
    const float* zb = Z + (size_t)b * h * w * ldz + c0;
    for (int i = threadIdx.x; i < rows * cols * 9 * LANES; i += 256) {
        const int c = i % LANES, pt = i / LANES;
        const int tap = pt % 9, px = pt / 9;
        const int ry = px / cols, rx = px - ry * cols;
        s[i] = __ldg(reinterpret_cast<const float4*>(zb + ((size_t)(y_lo + ry) * w + (x_lo + rx)) * ldz + tap * C) + c);
    }
    __syncthreads();
    const float slope = __ldg(prelu);
    const int cq = threadIdx.x % LANES;
    const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias + c0) + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
    const int row_step = cols * 9 * LANES;                          // float4 elements between footprint rows / columns
    constexpr int col_step = 9 * LANES;
#pragma unroll 1
    for (int pass = 0; pass < UF_T * UF_T / PIX_PER_PASS; ++pass) {
        const int pix = pass * PIX_PER_PASS + (threadIdx.x / LANES);
        const int Y = Y0 + pix / UF_T, X = X0 + pix % UF_T;
        if (Y >= H || X >= W) continue;
        // per axis and tap offset k - 1: footprint index of the first sample, step to the second one, the two weights (0 outside the map)
        int yo[3], ys[3], xo[3], xs[3];
        float wy0[3], wy1[3], wx0[3], wx1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int yy = Y + k - 1, xx = X + k - 1;
            const bool vy = yy >= 0 && yy < H, vx = xx >= 0 && xx < W;
            const float sy = rh * (vy ? yy : 0), sx = rw * (vx ? xx : 0);
            const int y0 = (int)sy, x0 = (int)sx;
            const float ly1 = sy - y0, lx1 = sx - x0;
            yo[k] = vy ? (y0 - y_lo) * row_step : 0; ys[k] = (vy && y0 < h - 1) ? row_step : 0;
            xo[k] = vx ? (x0 - x_lo) * col_step : 0; xs[k] = (vx && x0 < w - 1) ? col_step : 0;
            wy0[k] = vy ? 1.0f - ly1 : 0.f; wy1[k] = vy ? ly1 : 0.f;
            wx0[k] = vx ? 1.0f - lx1 : 0.f; wx1[k] = vx ? lx1 : 0.f;
        }
        float4 acc = bv;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                // (a tap outside the output map has zero weights and reads the first sample of the footprint)
                const float4* p = s + yo[ky] + xo[kx] + (ky * 3 + kx) * LANES + cq;
                const float4 v00 = p[0];
                const float4 v01 = p[xs[kx]];
                const float4 v10 = p[ys[ky]];
                const float4 v11 = p[ys[ky] + xs[kx]];
                const float w00 = wy0[ky] * wx0[kx], w01 = wy0[ky] * wx1[kx], w10 = wy1[ky] * wx0[kx], w11 = wy1[ky] * wx1[kx];
                acc.x = fmaf(w00, v00.x, acc.x); acc.y = fmaf(w00, v00.y, acc.y); acc.z = fmaf(w00, v00.z, acc.z); acc.w = fmaf(w00, v00.w, acc.w);
                acc.x = fmaf(w01, v01.x, acc.x); acc.y = fmaf(w01, v01.y, acc.y); acc.z = fmaf(w01, v01.z, acc.z); acc.w = fmaf(w01, v01.w, acc.w);
                acc.x = fmaf(w10, v10.x, acc.x); acc.y = fmaf(w10, v10.y, acc.y); acc.z = fmaf(w10, v10.z, acc.z); acc.w = fmaf(w10, v10.w, acc.w);
                acc.x = fmaf(w11, v11.x, acc.x); acc.y = fmaf(w11, v11.y, acc.y); acc.z = fmaf(w11, v11.z, acc.z); acc.w = fmaf(w11, v11.w, acc.w);
            }
        }
        acc.x = acc.x > 0.f ? acc.x : slope * acc.x; acc.y = acc.y > 0.f ? acc.y : slope * acc.y;
        acc.z = acc.z > 0.f ? acc.z : slope * acc.z; acc.w = acc.w > 0.f ? acc.w : slope * acc.w;
        *reinterpret_cast<float4*>(out + (((size_t)b * H + Y) * W + X) * ldo + c0 + cq * 4) = acc;
    }
}

// The same with TWO channel quads per thread (one pass of 64 pixels x 4 lanes instead of two passes of 32 x 8): the per-pixel index /
// weight arithmetic (~100 of the 355 instructions per output float4) is paid once per two outputs.  A quarter-warp then holds two
// pixels whose 64-byte halves of a 128-byte slot would collide on banks 0-15; the odd pixel of each pair takes the upper quads first, so
// every 8-thread phase reads banks 0-15 and 16-31 from one pixel each.  Each accumulator sees the same FFMA sequence as in the
// one-quad kernel: bit-identical results (scripts/upconv_ab.py hashes).
__global__ void __launch_bounds__(256)
upconv_finish_smem_fma2_kernel(const float* __restrict__ Z, int ldz, const float* __restrict__ bias, const float* __restrict__ prelu,
                               float* __restrict__ out, int ldo, int h, int w, int C, float rh, float rw, int tiles_x)
{
    constexpr int LANES = 8;
    extern __shared__ __align__(16) uint8_t uf_smem[];
    float4* s = reinterpret_cast<float4*>(uf_smem);                 // [(ry * cols + rx) * 9 + tap][LANES]
    const int H = 2 * h, W = 2 * w;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int Y0 = ty * UF_T, X0 = tx * UF_T, c0 = blockIdx.y * (LANES * 4), b = blockIdx.z;
    const int ya = max(Y0 - 1, 0), yb = min(Y0 + UF_T, H - 1), xa = max(X0 - 1, 0), xb = min(X0 + UF_T, W - 1);
    const int y_lo = (int)(rh * ya), x_lo = (int)(rw * xa);
    const int y_hi = min((int)(rh * yb) + 1, h - 1), x_hi = min((int)(rw * xb) + 1, w - 1);
    const int rows = y_hi - y_lo + 1, cols = x_hi - x_lo + 1;       // <= UF_FP each
    const float* zb = Z + (size_t)b * h * w * ldz + c0;
    for (int i = threadIdx.x; i < rows * cols * 9 * LANES; i += 256) {
        const int c = i % LANES, pt = i / LANES;
        const int tap = pt % 9, px = pt / 9;
        const int ry = px / cols, rx = px - ry * cols;
        s[i] = __ldg(reinterpret_cast<const float4*>(zb + ((size_t)(y_lo + ry) * w + (x_lo + rx)) * ldz + tap * C) + c);
    }
    __syncthreads();
    const float slope = __ldg(prelu);
    const int pix = threadIdx.x >> 2, l4 = threadIdx.x & 3;
    const int Y = Y0 + pix / UF_T, X = X0 + pix % UF_T;
    if (Y >= H || X >= W) return;
    const int first = (pix & 1) ? 4 : 0;
    const int qa = l4 + first, qb = l4 + (4 - first);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acca = bias ? __ldg(reinterpret_cast<const float4*>(bias + c0) + qa) : zero4;
    float4 accb = bias ? __ldg(reinterpret_cast<const float4*>(bias + c0) + qb) : zero4;
    const int row_step = cols * 9 * LANES;                          // float4 elements between footprint rows / columns
    constexpr int col_step = 9 * LANES;
    // per axis and tap offset k - 1: footprint index of the first sample, step to the second one, the two weights (0 outside the map)
    int yo[3], ys[3], xo[3], xs[3];
    float wy0[3], wy1[3], wx0[3], wx1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int yy = Y + k - 1, xx = X + k - 1;
        const bool vy = yy >= 0 && yy < H, vx = xx >= 0 && xx < W;
        const float sy = rh * (vy ? yy : 0), sx = rw * (vx ? xx : 0);
        const int y0 = (int)sy, x0 = (int)sx;
        const float ly1 = sy - y0, lx1 = sx - x0;
        yo[k] = vy ? (y0 - y_lo) * row_step : 0; ys[k] = (vy && y0 < h - 1) ? row_step : 0;
        xo[k] = vx ? (x0 - x_lo) * col_step : 0; xs[k] = (vx && x0 < w - 1) ? col_step : 0;
        wy0[k] = vy ? 1.0f - ly1 : 0.f; wy1[k] = vy ? ly1 : 0.f;
        wx0[k] = vx ? 1.0f - lx1 : 0.f; wx1[k] = vx ? lx1 : 0.f;
    }
#define DF_UF_FMA4(acc, wgt, v) acc.x = fmaf(wgt, v.x, acc.x); acc.y = fmaf(wgt, v.y, acc.y); acc.z = fmaf(wgt, v.z, acc.z); acc.w = fmaf(wgt, v.w, acc.w)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            // (a tap outside the output map has zero weights and reads the first sample of the footprint)
            const float4* p = s + yo[ky] + xo[kx] + (ky * 3 + kx) * LANES;
            const float4 a00 = p[qa], b00 = p[qb];
            const float4 a01 = p[xs[kx] + qa], b01 = p[xs[kx] + qb];
            const float4 a10 = p[ys[ky] + qa], b10 = p[ys[ky] + qb];
            const float4 a11 = p[ys[ky] + xs[kx] + qa], b11 = p[ys[ky] + xs[kx] + qb];
            const float w00 = wy0[ky] * wx0[kx], w01 = wy0[ky] * wx1[kx], w10 = wy1[ky] * wx0[kx], w11 = wy1[ky] * wx1[kx];
            DF_UF_FMA4(acca, w00, a00); DF_UF_FMA4(acca, w01, a01); DF_UF_FMA4(acca, w10, a10); DF_UF_FMA4(acca, w11, a11);
            DF_UF_FMA4(accb, w00, b00); DF_UF_FMA4(accb, w01, b01); DF_UF_FMA4(accb, w10, b10); DF_UF_FMA4(accb, w11, b11);
        }
    }
#undef DF_UF_FMA4
    acca.x = acca.x > 0.f ? acca.x : slope * acca.x; acca.y = acca.y > 0.f ? acca.y : slope * acca.y;
    acca.z = acca.z > 0.f ? acca.z : slope * acca.z; acca.w = acca.w > 0.f ? acca.w : slope * acca.w;
    accb.x = accb.x > 0.f ? accb.x : slope * accb.x; accb.y = accb.y > 0.f ? accb.y : slope * accb.y;
    accb.z = accb.z > 0.f ? accb.z : slope * accb.z; accb.w = accb.w > 0.f ? accb.w : slope * accb.w;
    float* o = out + (((size_t)b * H + Y) * W + X) * ldo + c0;
    *reinterpret_cast<float4*>(o + qa * 4) = acca;
    *reinterpret_cast<float4*>(o + qb * 4) = accb;
}

// (A separable form -- x pass at the footprint rows, then a y pass: 24 instead of 36 shared-memory reads per output -- was measured in
// round 2 and is SLOWER, 0.357 vs 0.195 ms on up_1: the extra barrier, the ragged x pass and the lower occupancy of a 78 KB CTA cost
// more than the reads saved; the kernel is latency / occupancy bound, not LDS bound.)
// Backward of upsample_nhwc_kernel in gather form (deterministic, no atomics): every INPUT pixel sums the output pixels
// whose 2x2 bilinear footprint contains it, with the forward's own index / weight arithmetic.
__global__ void __launch_bounds__(256)
upsample_nhwc_backward_kernel(const float* __restrict__ gout, int ldo, float* __restrict__ gin, int ldi, int B, int hin, int win,
                              int hout, int wout, int C, float rh, float rw, int align)
{
    const unsigned c4 = C >> 2;
    const unsigned total = (unsigned)B * hin * win * c4;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int cq = (int)(i % c4);
        unsigned r = i / c4;
        const int x = (int)(r % (unsigned)win); r /= (unsigned)win;
        const int y = (int)(r % (unsigned)hin), b = (int)(r / (unsigned)hin);
        // candidate output rows / columns: source coordinate within (y-1, y+1); conservative bounds, exact test inside
        int Y0 = 0, Y1 = hout - 1, X0 = 0, X1 = wout - 1;
        const float off = align ? 0.0f : 0.5f;           // source = r * (out + off) - off
        if (rh > 0.f) { Y0 = max(0, (int)floorf((y - 1 + off) / rh - off) - 1); Y1 = min(hout - 1, (int)ceilf((y + 1 + off) / rh - off) + 1); }
        if (rw > 0.f) { X0 = max(0, (int)floorf((x - 1 + off) / rw - off) - 1); X1 = min(wout - 1, (int)ceilf((x + 1 + off) / rw - off) + 1); }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int Y = Y0; Y <= Y1; ++Y) {
            float sy = align ? rh * Y : fmaxf(rh * (Y + 0.5f) - 0.5f, 0.f);
            const int y0 = (int)sy;
            const int yp = y0 < hin - 1 ? 1 : 0;
            const float ly1 = sy - y0, ly0 = 1.0f - ly1;
            const float wy = (y0 == y ? ly0 : 0.f) + (y0 + yp == y ? ly1 : 0.f);
            if (wy == 0.f) continue;
            for (int X = X0; X <= X1; ++X) {
                float sx = align ? rw * X : fmaxf(rw * (X + 0.5f) - 0.5f, 0.f);
                const int x0 = (int)sx;
                const int xp = x0 < win - 1 ? 1 : 0;
                const float lx1 = sx - x0, lx0 = 1.0f - lx1;
                const float wx = (x0 == x ? lx0 : 0.f) + (x0 + xp == x ? lx1 : 0.f);
                if (wx == 0.f) continue;
                const float4 g = __ldg(reinterpret_cast<const float4*>(gout + (((size_t)b * hout + Y) * wout + X) * ldo + cq * 4));
                const float wgt = wy * wx;
                acc.x += wgt * g.x; acc.y += wgt * g.y; acc.z += wgt * g.z; acc.w += wgt * g.w;
            }
        }
        *reinterpret_cast<float4*>(gin + (((size_t)b * hin + y) * win + x) * ldi + cq * 4) = acc;
    }
}

// log_softmax over the 32 channels of every pixel (lib/pspnet.py:53-56), one warp per pixel, in place
__global__ void __launch_bounds__(256)
log_softmax32_kernel(float* __restrict__ x, long long pixels)
{
    const int lane = threadIdx.x & 31;
    for (long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < pixels;
         p += (long long)gridDim.x * (blockDim.x >> 5)) {
        const float v = x[p * 32 + lane];
        float m = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float s = df::warp_sum(expf(v - m));
        x[p * 32 + lane] = (v - m) - logf(s);
    }
}

// Last decoder stage evaluated only where the head looks (the N `choose` pixels per crop instead of H*W):
// A[(b*N + n), tap*C + c] = U(b, y + dy, x + dx, c) for the 3x3 neighbourhood of pixel choose[b,n] = y*W + x, where
// U is the x2 bilinear (align_corners) up-sampling of `in` (B,h,w,C) to (H,W) = (2h,2w) and positions outside the
// image are the convolution's zero padding.  Same interpolation arithmetic as upsample_nhwc_kernel, so the gathered
// patches equal the dense tensor's values bit for bit.  One thread per (point, tap, 4 channels).
__global__ void __launch_bounds__(256)
gather_up_patches_kernel(const float* __restrict__ in, const int64_t* __restrict__ choose, float* __restrict__ A, int B, int N,
                         int h, int w, int C, float rh, float rw)
{
    const unsigned c4 = C >> 2;
    const unsigned total = (unsigned)B * N * 9 * c4;
    const int H = 2 * h, W = 2 * w;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int cq = (int)(i % c4);
        unsigned r = i / c4;
        const int tap = (int)(r % 9u);
        const unsigned pt = r / 9u;
        const int b = (int)(pt / (unsigned)N);
        long long pix = choose[pt];
        pix = pix < 0 ? 0 : (pix >= (long long)H * W ? (long long)H * W - 1 : pix);
        const int y = (int)(pix / W) + tap / 3 - 1, x = (int)(pix % W) + tap % 3 - 1;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const float sy = rh * y, sx = rw * x;
            const int y0 = (int)sy, x0 = (int)sx;
            const int yp = y0 < h - 1 ? 1 : 0, xp = x0 < w - 1 ? 1 : 0;
            const float ly1 = sy - y0, ly0 = 1.0f - ly1, lx1 = sx - x0, lx0 = 1.0f - lx1;
            const float* p = in + (((size_t)b * h + y0) * w + x0) * C + cq * 4;
            const float4 v00 = __ldg(reinterpret_cast<const float4*>(p));
            const float4 v01 = __ldg(reinterpret_cast<const float4*>(p + (size_t)xp * C));
            const float4 v10 = __ldg(reinterpret_cast<const float4*>(p + (size_t)yp * w * C));
            const float4 v11 = __ldg(reinterpret_cast<const float4*>(p + ((size_t)yp * w + xp) * C));
            o.x = ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
            o.y = ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
            o.z = ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
            o.w = ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
        }
        reinterpret_cast<float4*>(A)[i] = o;
    }
}

// gather_up_patches with one thread per (point, 4 channels) walking the nine taps: the 36 corner reads of a point touch at most nine
// low-resolution pixels, and read by ONE thread back to back all but the first touch of each are L1 hits (one thread per tap spreads them over
// 9 x 16 threads: measured 1.8 TB/s of patch bytes, the SM's L2 port carrying four bytes per byte written).  Same arithmetic.  Same box,
// 3 launches per step: 0.169-0.174 -> 0.124-0.146 ms (one eager step with the buckets' streams overlapping: 0.18 once).
__global__ void __launch_bounds__(256)
gather_up_patches_taps_kernel(const float* __restrict__ in, const int64_t* __restrict__ choose, float* __restrict__ A, int B, int N,
                              int h, int w, int C, float rh, float rw)
{
    const unsigned c4 = C >> 2;
    const unsigned total = (unsigned)B * N * c4;
    const int H = 2 * h, W = 2 * w;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int cq = (int)(i % c4);
        const unsigned pt = i / c4;
        const int b = (int)(pt / (unsigned)N);
        long long pix = choose[pt];
        pix = pix < 0 ? 0 : (pix >= (long long)H * W ? (long long)H * W - 1 : pix);
        const int yc = (int)(pix / W), xc = (int)(pix % W);
        float4* dst = reinterpret_cast<float4*>(A) + (size_t)pt * 9 * c4 + cq;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int y = yc + tap / 3 - 1, x = xc + tap % 3 - 1;
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y >= 0 && y < H && x >= 0 && x < W) {
                const float sy = rh * y, sx = rw * x;
                const int y0 = (int)sy, x0 = (int)sx;
                const int yp = y0 < h - 1 ? 1 : 0, xp = x0 < w - 1 ? 1 : 0;
                const float ly1 = sy - y0, ly0 = 1.0f - ly1, lx1 = sx - x0, lx0 = 1.0f - lx1;
                const float* p = in + (((size_t)b * h + y0) * w + x0) * C + cq * 4;
                const float4 v00 = __ldg(reinterpret_cast<const float4*>(p));
                const float4 v01 = __ldg(reinterpret_cast<const float4*>(p + (size_t)xp * C));
                const float4 v10 = __ldg(reinterpret_cast<const float4*>(p + (size_t)yp * w * C));
                const float4 v11 = __ldg(reinterpret_cast<const float4*>(p + ((size_t)yp * w + xp) * C));
                o.x = ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
                o.y = ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
                o.z = ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
                o.w = ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
            }
            dst[(size_t)tap * c4] = o;
        }
    }
}

// DF_ENC_V1=1: the first-generation pyramid / patch-gather kernels (A/B timing runs)
inline bool enc_v1()
{
    static const int v = getenv("DF_ENC_V1") ? atoi(getenv("DF_ENC_V1")) : 0;
    return v != 0;
}

inline unsigned grid_for(long long total, int per_block)
{
    long long b = (total + per_block - 1) / per_block;
    const long long cap = 148LL * 32;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" int df_enc_im2col_conv1(const float* img, float* A, int B, int H, int W, int ldk, void* stream)
{
    if (!img || !A || B <= 0 || H <= 0 || W <= 0 || ldk < 147 || ldk > 192 || (ldk & 3) || ((uintptr_t)A & 15)) return DF_ERR_ARG;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    if ((long long)B * Ho * Wo * ldk >= (1LL << 31)) return DF_ERR_ARG;
    im2col_conv1_kernel<<<grid_for((long long)B * Ho * Wo * (ldk >> 2), 256), 256, 0, (cudaStream_t)stream>>>(img, A, B, H, W, Ho, Wo, ldk);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_enc_maxpool(const float* in, float* out, int B, int H, int W, int C, void* stream)
{
    if (!in || !out || B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3)) return DF_ERR_ARG;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long total = (long long)B * Ho * Wo * (C >> 2);
    // (unsigned index arithmetic needs total + one grid stride < 2^32)
    if (total < (1LL << 31)) maxpool_kernel<unsigned><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, H, W, C, Ho, Wo);
    else maxpool_kernel<long long><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, H, W, C, Ho, Wo);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_enc_im2col_s2(const float* in, float* A, int B, int H, int W, int C, void* stream)
{
    if (!in || !A || B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3)) return DF_ERR_ARG;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long total = (long long)B * Ho * Wo * 9 * (C >> 2);
    if (total < (1LL << 31)) im2col_s2_kernel<unsigned><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, A, B, H, W, C, Ho, Wo);
    else im2col_s2_kernel<long long><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, A, B, H, W, C, Ho, Wo);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_enc_col2im_s2(const float* dA, float* dx, int B, int H, int W, int C, void* stream)
{
    if (!dA || !dx || B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3)) return DF_ERR_ARG;
    if (((uintptr_t)dA & 15) || ((uintptr_t)dx & 15)) return DF_ERR_ARG;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    col2im_s2_kernel<<<grid_for((long long)B * H * W * (C >> 2), 256), 256, 0, (cudaStream_t)stream>>>(dA, dx, B, H, W, C, Ho, Wo);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_enc_adaptive_avgpool(const float* in, int ldi, float* out, int B, int H, int W, int C, int S, void* stream)
{
    if (!in || !out || B <= 0 || H <= 0 || W <= 0 || C <= 0 || S <= 0 || ldi < C) return DF_ERR_ARG;
    adaptive_avgpool_kernel<<<grid_for((long long)B * S * S * C, 256), 256, 0, (cudaStream_t)stream>>>(in, ldi, out, B, H, W, C, S);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_enc_pyramid_pool(const float* in, int ldi, float* out, int B, int H, int W, int C, void* stream)
{
    if (!in || !out || B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3) || (ldi & 3) || ldi < C) return DF_ERR_ARG;
    if (((uintptr_t)in & 15) || ((uintptr_t)out & 15) || (long long)B * 50 * (C >> 2) >= (1LL << 31)) return DF_ERR_ARG;
    // DF_ENC_POOL_SPLIT: stages whose cells are shared out over a CTA (pyramid_pool_split_kernel): 0 none, 1 (default) the 1 x 1 and 2 x 2
    // stages (100+ pixels per cell), 2 all.  Measured on the bench shapes (3 launches per step, same box): none 0.129 ms, all 0.209 ms (a
    // CTA and two barriers for a 12-pixel cell of the 6 x 6 stage), 1 x 1 and 2 x 2 only 0.055 + 0.058 ms, profiles/r2_s5_small_kernels.txt
    static const int split = getenv("DF_ENC_POOL_SPLIT") ? atoi(getenv("DF_ENC_POOL_SPLIT")) : 1;
    const unsigned rows = enc_v1() || split == 0 ? 0u : (split == 1 ? 5u * B : 50u * B);
    const int groups = ((C >> 2) + 31) >> 5;
    if (rows) pyramid_pool_split_kernel<<<grid_for((long long)rows * groups, 1), 256, 0, (cudaStream_t)stream>>>(in, ldi, out, B, H, W, C, rows);
    if (rows < 50u * B)
        pyramid_pool_kernel<<<grid_for((long long)(50u * B - rows) * (C >> 2), 256), 256, 0, (cudaStream_t)stream>>>(in, ldi, out, B, H, W, C, rows);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_enc_pyramid_sum(const float* Y, float* out, int ldo, int B, int H, int W, int C, void* stream)
{
    if (!Y || !out || B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3) || (ldo & 3) || ldo < C) return DF_ERR_ARG;
    if (((uintptr_t)Y & 15) || ((uintptr_t)out & 15) || (long long)B * H * W * (C >> 2) >= (1LL << 31)) return DF_ERR_ARG;
    if (enc_v1()) pyramid_sum_kernel<<<grid_for((long long)B * H * W * (C >> 2), 256), 256, 0, (cudaStream_t)stream>>>(Y, out, ldo, B, H, W, C);
    else {
        // DF_ENC_SUM: 1 = rows kernel (corner vectors from L2), default = the same walk over cells staged in shared memory
        static const int form = getenv("DF_ENC_SUM") ? atoi(getenv("DF_ENC_SUM")) : 2;
        if (form == 1) pyramid_sum_rows_kernel<<<grid_for((long long)B * H * (C >> 2), 256), 256, 0, (cudaStream_t)stream>>>(Y, out, ldo, B, H, W, C);
        else pyramid_sum_smem_kernel<<<grid_for((long long)B * (((C >> 2) + 31) >> 5), 1), PS_WARPS * 32, 0, (cudaStream_t)stream>>>(Y, out, ldo, B, H, W, C);
    }
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_enc_upconv_finish(const float* Z, int ldz, const float* bias, const float* prelu, float* out, int ldo, int B,
                                    int h, int w, int C, void* stream)
{
    if (!Z || !prelu || !out || B <= 0 || h <= 0 || w <= 0 || C <= 0 || (C & 3) || (ldz & 3) || (ldo & 3) || ldz < 9 * C || ldo < C)
        return DF_ERR_ARG;
    if (((uintptr_t)Z & 15) || ((uintptr_t)out & 15) || ((uintptr_t)bias & 15)) return DF_ERR_ARG;
    if ((long long)B * 4 * h * w * (C >> 2) >= (1LL << 31)) return DF_ERR_ARG;
    const int H = 2 * h, W = 2 * w;
    const float rh = H > 1 ? (float)(h - 1) / (H - 1) : 0.f, rw = W > 1 ? (float)(w - 1) / (W - 1) : 0.f;
    // DF_UPCONV_SMEM: 0 = L1-gather kernel, 1 = footprint staged in shared memory, 32 channels per CTA (bit-identical to 0), 2 = the same
    // with 16 channels per CTA (twice the CTAs per SM; measured 10% SLOWER: occupancy is not what limits this kernel), 3 (default) = 1
    // with per-axis quantities hoisted and one FFMA per corner and channel (355 instead of 524 SASS instructions per output float4,
    // branch-free): 0.1956 -> 0.1748 ms on up_1's 64x20x20x256, -10..11% on every bench shape (profiles/r2_s4_upconv_ab.jsonl)
    // 4 = 3 with two channel quads per thread (18% fewer instructions, bit-identical): 0.174 -> 0.170 ms, +2.5% on the smallest bench shape
    // (profiles/r2_s5_upconv4_ab.jsonl) -- LDS and FFMA time are balanced in this kernel, the instruction count is not what paces it
    static const int use_smem = getenv("DF_UPCONV_SMEM") ? atoi(getenv("DF_UPCONV_SMEM")) : 3;
    if (use_smem && C % 32 == 0 && B <= 65535 && h >= 2 && w >= 2) {
        static bool attr_done = false;
        if (!attr_done) {
            cudaError_t e = cudaFuncSetAttribute(upconv_finish_smem_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, UF_SMEM);
            if (e != cudaSuccess) return (int)e;
            attr_done = true;
        }
        const int tiles_x = (W + UF_T - 1) / UF_T, tiles_y = (H + UF_T - 1) / UF_T;
        if (use_smem == 4) {
            static bool attr4 = false;
            if (!attr4) {
                cudaError_t e = cudaFuncSetAttribute(upconv_finish_smem_fma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UF_SMEM);
                if (e != cudaSuccess) return (int)e;
                attr4 = true;
            }
            upconv_finish_smem_fma2_kernel<<<dim3(tiles_x * tiles_y, C / 32, B), 256, UF_SMEM, (cudaStream_t)stream>>>(Z, ldz, bias, prelu, out, ldo, h, w, C, rh, rw, tiles_x);
        } else if (use_smem == 3) {
            static bool attr3 = false;
            if (!attr3) {
                cudaError_t e = cudaFuncSetAttribute(upconv_finish_smem_fma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, UF_SMEM);
                if (e != cudaSuccess) return (int)e;
                attr3 = true;
            }
            upconv_finish_smem_fma_kernel<8><<<dim3(tiles_x * tiles_y, C / 32, B), 256, UF_SMEM, (cudaStream_t)stream>>>(Z, ldz, bias, prelu, out, ldo, h, w, C, rh, rw, tiles_x);
        } else if (use_smem == 2)
            upconv_finish_smem_kernel<4><<<dim3(tiles_x * tiles_y, C / 16, B), 256, UF_SMEM / 2, (cudaStream_t)stream>>>(Z, ldz, bias, prelu, out, ldo, h, w, C, rh, rw, tiles_x);
        else
            upconv_finish_smem_kernel<8><<<dim3(tiles_x * tiles_y, C / 32, B), 256, UF_SMEM, (cudaStream_t)stream>>>(Z, ldz, bias, prelu, out, ldo, h, w, C, rh, rw, tiles_x);
        DF_RETURN_LAST_ERROR();
    }
    upconv_finish_kernel<<<grid_for((long long)B * H * W * (C >> 2), 256), 256, 0, (cudaStream_t)stream>>>(
        Z, ldz, bias, prelu, out, ldo, B, h, w, C, rh, rw);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_enc_upsample(const float* in, int ldi, float* out, int ldo, int B, int hin, int win, int hout, int wout,
                               int C, int align_corners, void* stream)
{
    if (!in || !out || B <= 0 || hin <= 0 || win <= 0 || hout <= 0 || wout <= 0 || C <= 0 || (C & 3) || (ldi & 3) || (ldo & 3))
        return DF_ERR_ARG;
    if (((uintptr_t)in & 15) || ((uintptr_t)out & 15)) return DF_ERR_ARG;
    if ((long long)B * hout * wout * (C >> 2) >= (1LL << 31)) return DF_ERR_ARG;
    float rh, rw;
    if (align_corners) {
        rh = hout > 1 ? (float)(hin - 1) / (hout - 1) : 0.f;
        rw = wout > 1 ? (float)(win - 1) / (wout - 1) : 0.f;
    } else {
        rh = (float)hin / hout;
        rw = (float)win / wout;
    }
    upsample_nhwc_kernel<<<grid_for((long long)B * hout * wout * (C >> 2), 256), 256, 0, (cudaStream_t)stream>>>(
        in, ldi, out, ldo, B, hin, win, hout, wout, C, rh, rw, align_corners);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_enc_log_softmax32(float* x, long long pixels, void* stream)
{
    if (!x || pixels <= 0) return DF_ERR_ARG;
    log_softmax32_kernel<<<grid_for(pixels, 8), 256, 0, (cudaStream_t)stream>>>(x, pixels);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_enc_gather_up_patches(const float* in, const int64_t* choose, float* A, int B, int N, int h, int w, int C,
                                        void* stream)
{
    if (!in || !choose || !A || B <= 0 || N <= 0 || h <= 0 || w <= 0 || C <= 0 || (C & 3)) return DF_ERR_ARG;
    if ((long long)B * N * 9 * (C >> 2) >= (1LL << 31)) return DF_ERR_ARG;
    const int H = 2 * h, W = 2 * w;
    const float rh = H > 1 ? (float)(h - 1) / (H - 1) : 0.f, rw = W > 1 ? (float)(w - 1) / (W - 1) : 0.f;
    if (enc_v1()) gather_up_patches_kernel<<<grid_for((long long)B * N * 9 * (C >> 2), 256), 256, 0, (cudaStream_t)stream>>>(
        in, choose, A, B, N, h, w, C, rh, rw);
    else gather_up_patches_taps_kernel<<<grid_for((long long)B * N * (C >> 2), 256), 256, 0, (cudaStream_t)stream>>>(
        in, choose, A, B, N, h, w, C, rh, rw);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_enc_upsample_backward(const float* gout, int ldo, float* gin, int ldi, int B, int hin, int win, int hout,
                                        int wout, int C, int align_corners, void* stream)
{
    if (!gout || !gin || B <= 0 || hin <= 0 || win <= 0 || hout <= 0 || wout <= 0 || C <= 0 || (C & 3) || (ldi & 3) || (ldo & 3))
        return DF_ERR_ARG;
    if (((uintptr_t)gout & 15) || ((uintptr_t)gin & 15)) return DF_ERR_ARG;
    if ((long long)B * hin * win * (C >> 2) >= (1LL << 31)) return DF_ERR_ARG;
    float rh, rw;
    if (align_corners) {
        rh = hout > 1 ? (float)(hin - 1) / (hout - 1) : 0.f;
        rw = wout > 1 ? (float)(win - 1) / (wout - 1) : 0.f;
    } else {
        rh = (float)hin / hout;
        rw = (float)win / wout;
    }
    upsample_nhwc_backward_kernel<<<grid_for((long long)B * hin * win * (C >> 2), 256), 256, 0, (cudaStream_t)stream>>>(
        gout, ldo, gin, ldi, B, hin, win, hout, wout, C, rh, rw, align_corners);
    DF_RETURN_LAST_ERROR();
}
