"""Autograd functions of the training encoder's element-wise / pooling ops on the kernels of csrc/train_ew.cu.

The module graph of lib/extractors.py / lib/pspnet.py (reference: lib/extractors.py:78-124, lib/pspnet.py:7-77) keeps its shape
and parameter names; in training on CUDA its non-convolution ops route through these functions instead of ATen:
3x3/2 max pooling, the four adaptive average pools (one pass), the pyramid concat (written slice by slice: four bilinear
resizes + one pitched copy), PReLU, Dropout2d (device-side counter-based mask), channel log-softmax.  ReLU and the skip
connections are epilogues of the convolutions (lib/conv_tc.py).  All tensors are channels_last fp32."""
from __future__ import annotations

import torch

from .._C import check, lib, ptr, stream

CL = torch.channels_last
ENABLED = True          # False: the module graph falls back to the ATen ops (A/B runs, strict-fp32 parity mode)


def usable(x: torch.Tensor) -> bool:
    from . import conv_tc
    return (ENABLED and conv_tc.ENABLED and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] % 4 == 0
            and torch.is_grad_enabled())


def nhwc(x: torch.Tensor) -> torch.Tensor:
    """fp32 channels_last storage (a no-op for the tensors our own kernels produce)."""
    x = x.float() if x.dtype != torch.float32 else x
    return x if x.is_contiguous(memory_format=CL) else x.contiguous(memory_format=CL)


def _empty_cl(b, c, h, w, device):
    return torch.empty(b, c, h, w, device=device, dtype=torch.float32, memory_format=CL)


def relu_mask(dy: torch.Tensor, act: torch.Tensor) -> torch.Tensor:
    """dy * [act > 0] for two channels_last tensors of the same shape (backward of a ReLU fused into a convolution epilogue)."""
    dy = nhwc(dy)
    b, c, h, w = dy.shape
    out = torch.empty_like(dy, memory_format=CL)
    check(lib.df_ew_relu_mask(ptr(dy), ptr(act), ptr(out), c, c, b * h * w, stream()), "df_ew_relu_mask")
    return out


def colsum(dy: torch.Tensor) -> torch.Tensor:
    """Bias gradient: sum over batch and pixels of a channels_last (B,C,H,W) gradient."""
    dy = nhwc(dy)
    b, c, h, w = dy.shape
    rows = b * h * w
    out = torch.empty(c, device=dy.device, dtype=torch.float32)
    # df_colsum_rows gives one block per (32 columns, group): cut the rows into groups so that ~4 blocks per SM are busy, then add the
    # per-group sums in fixed order (deterministic)
    groups = max(1, min(rows // 512, 592 // max(1, (c + 31) // 32)))
    if groups == 1:
        check(lib.df_colsum_rows(ptr(dy), c, rows, 1, c, ptr(out), 0, stream()), "df_colsum_rows")
        return out
    rpg = rows // groups
    tail = rows - rpg * groups
    part = torch.empty(groups + (1 if tail else 0), c, device=dy.device, dtype=torch.float32)
    check(lib.df_colsum_rows(ptr(dy), c, rpg, groups, c, ptr(part), 0, stream()), "df_colsum_rows")
    if tail:
        check(lib.df_colsum_rows(ptr(dy) + rpg * groups * c * 4, c, tail, 1, c, ptr(part) + groups * c * 4, 0, stream()), "df_colsum_rows")
    check(lib.df_reduce_partials(ptr(part), part.shape[0], c, ptr(out), 0, stream()), "df_reduce_partials")
    return out


class MaxPoolFn(torch.autograd.Function):
    """F.max_pool2d(x, 3, stride=2, padding=1)"""

    @staticmethod
    def forward(ctx, x):
        xn = nhwc(x.detach())
        b, c, h, w = xn.shape
        y = _empty_cl(b, c, (h - 1) // 2 + 1, (w - 1) // 2 + 1, xn.device)
        check(lib.df_enc_maxpool(ptr(xn), ptr(y), b, h, w, c, stream()), "df_enc_maxpool")
        ctx.save_for_backward(xn)
        return y

    @staticmethod
    def backward(ctx, g):
        (xn,) = ctx.saved_tensors
        b, c, h, w = xn.shape
        g = nhwc(g)
        gi = torch.empty_like(xn, memory_format=CL)
        check(lib.df_ew_maxpool_backward(ptr(xn), ptr(g), ptr(gi), b, h, w, c, stream()), "df_ew_maxpool_backward")
        return gi


_SIZES = (1, 2, 3, 6)


class PyramidPoolFn(torch.autograd.Function):
    """The four nn.AdaptiveAvgPool2d((s, s)), s = 1, 2, 3, 6, of lib/pspnet.py:10-15 in one pass; returns the four pooled maps."""

    @staticmethod
    def forward(ctx, feats):
        f = nhwc(feats.detach())
        b, c, h, w = f.shape
        pooled = torch.empty(50 * b, c, device=f.device, dtype=torch.float32)
        check(lib.df_enc_pyramid_pool(ptr(f), c, ptr(pooled), b, h, w, c, stream()), "df_enc_pyramid_pool")
        ctx.shape = (b, c, h, w)
        outs, row = [], 0
        for s in _SIZES:
            outs.append(pooled[row:row + b * s * s].view(b, s, s, c).permute(0, 3, 1, 2))      # channels_last view of the stage block
            row += b * s * s
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        b, c, h, w = ctx.shape
        dev = next(g for g in grads if g is not None).device
        dpool = torch.empty(50 * b, c, device=dev, dtype=torch.float32)
        row = 0
        for s, g in zip(_SIZES, grads):
            rows = b * s * s
            if g is None:
                dpool[row:row + rows].zero_()
            else:
                g = g.float()
                gm = g.permute(0, 2, 3, 1)
                if not gm.is_contiguous():
                    gm = gm.contiguous()
                check(lib.df_ew_copy2d(ptr(gm), c, ptr(dpool) + row * c * 4, c, rows, c, stream()), "df_ew_copy2d")
            row += rows
        dx = _empty_cl(b, c, h, w, dev)
        check(lib.df_ew_pyramid_pool_backward(ptr(dpool), ptr(dx), c, b, h, w, c, stream()), "df_ew_pyramid_pool_backward")
        return dx


class PyramidCatFn(torch.autograd.Function):
    """torch.cat([upsample(y_s, (H, W), bilinear, align_corners=False) for s in 1,2,3,6] + [feats], 1) (lib/pspnet.py:20-23): the
    four resizes write straight into their channel slice of the concat buffer, `feats` is a pitched copy -- no separate resized maps."""

    @staticmethod
    def forward(ctx, feats, *ys):
        f = nhwc(feats.detach())
        b, c, h, w = f.shape
        ys = [nhwc(y.detach()) for y in ys]
        ctot = c + sum(y.shape[1] for y in ys)
        out = _empty_cl(b, ctot, h, w, f.device)
        off = 0
        for y in ys:
            cy, s = y.shape[1], y.shape[2]
            check(lib.df_enc_upsample(ptr(y), cy, ptr(out) + off * 4, ctot, b, s, y.shape[3], h, w, cy, 0, stream()), "df_enc_upsample")
            off += cy
        check(lib.df_ew_copy2d(ptr(f), c, ptr(out) + off * 4, ctot, b * h * w, c, stream()), "df_ew_copy2d")
        ctx.meta = (b, c, h, w, ctot, [(y.shape[1], y.shape[2], y.shape[3]) for y in ys])
        return out

    @staticmethod
    def backward(ctx, g):
        b, c, h, w, ctot, ymeta = ctx.meta
        g = nhwc(g)
        gys, off = [], 0
        for cy, sh, sw in ymeta:
            gy = _empty_cl(b, cy, sh, sw, g.device)
            check(lib.df_enc_upsample_backward(ptr(g) + off * 4, ctot, ptr(gy), cy, b, sh, sw, h, w, cy, 0, stream()),
                  "df_enc_upsample_backward")
            gys.append(gy)
            off += cy
        gf = _empty_cl(b, c, h, w, g.device)
        check(lib.df_ew_copy2d(ptr(g) + off * 4, ctot, ptr(gf), c, b * h * w, c, stream()), "df_ew_copy2d")
        return (gf, *gys)


class PReLUFn(torch.autograd.Function):
    """nn.PReLU() with its single slope"""

    @staticmethod
    def forward(ctx, x, slope):
        xn = nhwc(x.detach())
        a = slope.detach().float().contiguous()
        y = torch.empty_like(xn, memory_format=CL)
        check(lib.df_ew_prelu(ptr(xn), ptr(a), ptr(y), xn.numel(), stream()), "df_ew_prelu")
        ctx.save_for_backward(xn, a)
        return y

    @staticmethod
    def backward(ctx, g):
        xn, a = ctx.saved_tensors
        g = nhwc(g)
        dx = torch.empty_like(xn, memory_format=CL)
        da = torch.empty(1, device=xn.device, dtype=torch.float32)
        scratch = torch.empty(int(lib.df_ew_prelu_scratch_floats()), device=xn.device, dtype=torch.float32)
        check(lib.df_ew_prelu_backward(ptr(xn), ptr(a), ptr(g), ptr(dx), ptr(da), ptr(scratch), xn.numel(), stream()), "df_ew_prelu_backward")
        return dx, da


_DROPOUT_STATE = {}


def dropout_state(device) -> torch.Tensor:
    """{seed, counter} of the Dropout2d masks on `device` (two int64 on the device; the mask kernel advances the counter itself, so a
    captured CUDA graph draws new masks at every replay).  Seeded from torch's default generator."""
    st = _DROPOUT_STATE.get(device)
    if st is None:
        st = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=device)
        _DROPOUT_STATE[device] = st
    return st


class Dropout2dFn(torch.autograd.Function):
    """nn.Dropout2d(p) in training mode: whole (sample, channel) maps are zeroed with probability p, the rest scaled by 1 / (1 - p)."""

    @staticmethod
    def forward(ctx, x, p):
        xn = nhwc(x.detach())
        b, c, h, w = xn.shape
        mask = torch.empty(b * c, device=xn.device, dtype=torch.float32)
        check(lib.df_ew_dropout_mask(ptr(mask), b * c, float(p), ptr(dropout_state(xn.device)), stream()), "df_ew_dropout_mask")
        y = torch.empty_like(xn, memory_format=CL)
        check(lib.df_ew_scale_bc(ptr(xn), ptr(mask), ptr(y), b, h * w, c, stream()), "df_ew_scale_bc")
        ctx.save_for_backward(mask)
        return y

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        g = nhwc(g)
        b, c, h, w = g.shape
        gi = torch.empty_like(g, memory_format=CL)
        check(lib.df_ew_scale_bc(ptr(g), ptr(mask), ptr(gi), b, h * w, c, stream()), "df_ew_scale_bc")
        return gi, None


class LogSoftmax32Fn(torch.autograd.Function):
    """nn.LogSoftmax(dim=1) over the 32 embedding channels"""

    @staticmethod
    def forward(ctx, x):
        xn = nhwc(x.detach())
        b, c, h, w = xn.shape
        y = torch.empty_like(xn, memory_format=CL)
        check(lib.df_ew_copy2d(ptr(xn), c, ptr(y), c, b * h * w, c, stream()), "df_ew_copy2d")
        check(lib.df_enc_log_softmax32(ptr(y), b * h * w, stream()), "df_enc_log_softmax32")
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        g = nhwc(g)
        b, c, h, w = y.shape
        gi = torch.empty_like(y, memory_format=CL)
        check(lib.df_ew_log_softmax32_backward(ptr(y), ptr(g), ptr(gi), b * h * w, stream()), "df_ew_log_softmax32_backward")
        return gi
