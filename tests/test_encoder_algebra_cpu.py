"""CPU-only, float64: the algebraic identities the tensor-core encoder and the weight-gradient GEMM rest on (DESIGN.md
section 4, "Encoder algebra" / section 5) hold exactly -- independent of any kernel.  Reference semantics:
lib/pspnet.py:17-24 (pyramid), :27-37 (decoder stage), torch conv2d weight gradient."""
import torch
import torch.nn.functional as F


def test_folded_pyramid_equals_concat_bottleneck():
    g = torch.Generator().manual_seed(0)
    B, C, H, W, Co = 2, 16, 10, 15, 24
    f = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    stages = [torch.randn(C, C, 1, 1, generator=g, dtype=torch.float64) for _ in range(4)]
    wb = torch.randn(Co, 5 * C, 1, 1, generator=g, dtype=torch.float64)
    bias = torch.randn(Co, generator=g, dtype=torch.float64)
    priors = [F.interpolate(F.conv2d(F.adaptive_avg_pool2d(f, (s, s)), w), size=(H, W), mode="bilinear", align_corners=False)
              for s, w in zip((1, 2, 3, 6), stages)]
    want = F.relu(F.conv2d(torch.cat(priors + [f], 1), wb, bias))
    # folded: (bottleneck slice x stage weight) at the pooled resolution, resized, summed; only f goes through K = C at full size
    acc = F.conv2d(f, wb[:, 4 * C:], bias)
    for i, s in enumerate((1, 2, 3, 6)):
        wf = (wb[:, i * C:(i + 1) * C, 0, 0] @ stages[i][:, :, 0, 0])[:, :, None, None]
        acc = acc + F.interpolate(F.conv2d(F.adaptive_avg_pool2d(f, (s, s)), wf), size=(H, W), mode="bilinear", align_corners=False)
    assert float((F.relu(acc) - want).abs().max()) < 1e-12


def test_decoder_stage_at_low_resolution_equals_conv_after_resize():
    g = torch.Generator().manual_seed(1)
    B, Ci, Co, h, w = 2, 8, 6, 5, 7
    x = torch.randn(B, Ci, h, w, generator=g, dtype=torch.float64)
    wt = torch.randn(Co, Ci, 3, 3, generator=g, dtype=torch.float64)
    bias = torch.randn(Co, generator=g, dtype=torch.float64)
    up = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    want = F.conv2d(up, wt, bias, padding=1)
    # Z = the nine 1x1 products at the low resolution; each is resized, shifted by its tap (zero outside the map) and summed
    H, W = 2 * h, 2 * w
    acc = bias.view(1, Co, 1, 1).expand(B, Co, H, W).clone()
    for ky in range(3):
        for kx in range(3):
            z = F.interpolate(F.conv2d(x, wt[:, :, ky:ky + 1, kx:kx + 1]), scale_factor=2, mode="bilinear", align_corners=True)
            zp = F.pad(z, (1, 1, 1, 1))
            acc += zp[:, :, ky:ky + H, kx:kx + W]
    assert float((acc - want).abs().max()) < 1e-12


def _tap_reach(extent, t, dil):
    """Mirror of the launcher's patch cost (gemm_tc.cu, df_conv_tc): taps (of 3) each patch of t pixels along one axis reaches."""
    total = 0
    for o in range(0, extent, t):
        hi = min(o + t, extent) - 1
        total += sum(1 for k in (-1, 0, 1) if hi + k * dil >= 0 and o + k * dil < extent)
    return total


def test_skipped_taps_only_see_padding():
    """A tap the kernel skips for a patch (q_tap_mask) contributes nothing: zeroing exactly those (patch, tap) pairs leaves
    the dilated convolution unchanged; and small patches skip more (what the patch chooser minimises)."""
    g = torch.Generator().manual_seed(2)
    H = W = 10
    dil, t = 4, 2
    x = torch.randn(1, 3, H, W, generator=g, dtype=torch.float64)
    wt = torch.randn(4, 3, 3, 3, generator=g, dtype=torch.float64)
    want = F.conv2d(x, wt, padding=dil, dilation=dil)
    got = torch.zeros_like(want)
    visits = 0
    for y0 in range(0, H, t):
        for x0 in range(0, W, t):
            for ky in range(3):
                for kx in range(3):
                    dy, dx = (ky - 1) * dil, (kx - 1) * dil
                    yh, xh = min(y0 + t, H) - 1, min(x0 + t, W) - 1
                    if not (yh + dy >= 0 and y0 + dy < H and xh + dx >= 0 and x0 + dx < W):
                        continue                                       # skipped: the shifted patch lies entirely in the padding
                    visits += 1
                    xp = F.pad(x, (dil, dil, dil, dil))
                    patch = xp[:, :, y0 + dy + dil:y0 + t + dy + dil, x0 + dx + dil:x0 + t + dx + dil]
                    got[:, :, y0:y0 + t, x0:x0 + t] += torch.einsum("bchw,oc->bohw", patch, wt[:, :, ky, kx])
    assert float((got - want).abs().max()) < 1e-12
    assert visits == _tap_reach(H, t, dil) * _tap_reach(W, t, dil) == 121          # of 25 patches x 9 taps = 225
    assert _tap_reach(10, 10, 4) == 3 and _tap_reach(10, 2, 1) == 15 and _tap_reach(10, 1, 1) == 28 and _tap_reach(20, 4, 4) == 13


def test_weight_gradient_as_gemm_over_padded_flat_pixels():
    """dW[co,tap,ci] = dY_T (Cout x P') . shifted X_T (Cin x P')^T with P' the zero-padded (row length a multiple of 4),
    flattened pixel axis: vertical tap offsets are k offsets of multiples of 4, horizontal ones are pre-shifted planes."""
    g = torch.Generator().manual_seed(3)
    B, H, W, Ci, Co, d = 2, 5, 7, 4, 3, 2
    x = torch.randn(B, Ci, H, W, generator=g, dtype=torch.float64)
    dy = torch.randn(B, Co, H, W, generator=g, dtype=torch.float64)
    w = torch.zeros(Co, Ci, 3, 3, dtype=torch.float64, requires_grad=True)
    (F.conv2d(x, w, padding=d, dilation=d) * dy).sum().backward()
    Hp, Wp = H + 2 * d, (W + 2 * d + 3) // 4 * 4
    P = B * Hp * Wp

    def cmajor(t, shift):
        flat = F.pad(t, (d, Wp - W - d, d, d)).permute(1, 0, 2, 3).reshape(t.shape[1], -1)
        out = torch.zeros_like(flat)
        lo, hi = max(0, -shift), min(P, P - shift)
        out[:, lo:hi] = flat[:, lo + shift:hi + shift]
        return out

    dyT = cmajor(dy, 0)
    planes = [cmajor(x, (kx - 1) * d) for kx in range(3)]
    got = torch.zeros(Co, Ci, 3, 3, dtype=torch.float64)
    for ky in range(3):
        k0 = (ky - 1) * d * Wp
        assert k0 % 4 == 0
        for kx in range(3):
            xs = torch.zeros_like(planes[kx])
            lo, hi = max(0, -k0), min(P, P - k0)
            xs[:, lo:hi] = planes[kx][:, lo + k0:hi + k0]                 # the TMA box starts k0 further, zero-filled outside
            got[:, :, ky, kx] = dyT @ xs.t()
    assert float((got - w.grad).abs().max()) < 1e-12
