// Bilinear up-sampling for the colour encoder's pyramid / decoder stages (reference: lib/pspnet.py:20-23,
// F.upsample(..., mode='bilinear') and :30-34, nn.Upsample(scale_factor=2, align_corners=True)).
//
// torch's NCHW kernel assigns one thread per OUTPUT PIXEL and loops over batch x channels inside the thread;
// with the encoder's small maps (10x10 .. 160x160) and 64..1024 channels x hundreds of crops that is a few
// hundred threads doing all the work -- it was 64% of the whole pose step (profiles/r1_call4_launch_list_step.json).
// Here every output element is its own thread (x fastest -> coalesced stores, the four taps hit L1/L2).
// Index arithmetic follows ATen's area_pixel_compute_source_index so results match torch's to fp32 rounding.
#include "df_common.cuh"
#include "../../include/densefusion_b200.h"

namespace {

__global__ void __launch_bounds__(256)
upsample_bilinear_kernel(const float* __restrict__ in, float* __restrict__ out, long long planes, int hin, int win,
                         int hout, int wout, float rh, float rw, int align)
{
    const long long total = planes * hout * wout;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % wout);
        const long long t = i / wout;
        const int y = (int)(t % hout);
        const long long pl = t / hout;
        float sy, sx;
        if (align) { sy = rh * y; sx = rw * x; }
        else {
            sy = rh * (y + 0.5f) - 0.5f; sy = sy < 0.f ? 0.f : sy;
            sx = rw * (x + 0.5f) - 0.5f; sx = sx < 0.f ? 0.f : sx;
        }
        const int y0 = (int)sy, x0 = (int)sx;
        const int yp = y0 < hin - 1 ? 1 : 0, xp = x0 < win - 1 ? 1 : 0;
        const float ly1 = sy - y0, ly0 = 1.0f - ly1, lx1 = sx - x0, lx0 = 1.0f - lx1;
        const float* p = in + pl * (long long)hin * win + (long long)y0 * win + x0;
        const float v00 = __ldg(p), v01 = __ldg(p + xp), v10 = __ldg(p + yp * win), v11 = __ldg(p + yp * win + xp);
        out[i] = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
    }
}

}  // namespace

extern "C" int df_upsample_bilinear(const float* in, float* out, long long planes, int hin, int win, int hout, int wout,
                                    int align_corners, void* stream)
{
    if (!in || !out || planes <= 0 || hin <= 0 || win <= 0 || hout <= 0 || wout <= 0) return DF_ERR_ARG;
    float rh, rw;
    if (align_corners) {
        rh = hout > 1 ? (float)(hin - 1) / (hout - 1) : 0.f;
        rw = wout > 1 ? (float)(win - 1) / (wout - 1) : 0.f;
    } else {
        rh = (float)hin / hout;
        rw = (float)win / wout;
    }
    const long long total = planes * hout * wout;
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 64) blocks = 148LL * 64;
    upsample_bilinear_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, planes, hin, win, hout, wout,
                                                                                rh, rw, align_corners);
    DF_RETURN_LAST_ERROR();
}
