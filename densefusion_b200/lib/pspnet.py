"""Pyramid-pooling decoder of the colour encoder.  Parameter names / shapes follow lib/pspnet.py:7-77 of the reference.
Inference runs on densefusion_b200.encoder; this module graph is the TRAINING path -- convolutions on lib/conv_tc.py, pooling /
concat / PReLU / Dropout2d / log-softmax on lib/ew.py (own kernels), torch ops only as the CPU / strict-fp32 fallback."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import extractors
from .conv_tc import conv2d


class _UpsampleFn(torch.autograd.Function):
    """df_upsample_bilinear forward; the encoder is trained through torch/cuDNN, so its backward stays the library's
    own aten kernel (same one F.interpolate's autograd would call)."""

    @staticmethod
    def forward(ctx, x, size, align_corners):
        from .. import ops
        ctx.in_shape, ctx.size, ctx.align = tuple(x.shape), (int(size[0]), int(size[1])), bool(align_corners)
        if x.dim() == 4 and x.shape[1] % 4 == 0 and x.shape[1] > 1 and x.is_contiguous(memory_format=torch.channels_last):
            # channels_last activations (tensor-core training path): resize in the NHWC storage, keep the layout
            from .._C import check, lib, ptr, stream
            b, c, h, w = x.shape
            out = torch.empty(b, c, ctx.size[0], ctx.size[1], device=x.device, dtype=torch.float32,
                              memory_format=torch.channels_last)
            check(lib.df_enc_upsample(ptr(x), c, ptr(out), c, b, h, w, ctx.size[0], ctx.size[1], c, 1 if ctx.align else 0,
                                      stream()), "df_enc_upsample")
            return out
        return ops.upsample_bilinear(x, size, align_corners)

    @staticmethod
    def backward(ctx, g):
        b, c, h, w = ctx.in_shape
        if c % 4 == 0 and c > 1 and g.is_cuda:
            from .._C import check, lib, ptr, stream
            g = g.float().contiguous(memory_format=torch.channels_last)
            gi = torch.empty(b, c, h, w, device=g.device, dtype=torch.float32, memory_format=torch.channels_last)
            check(lib.df_enc_upsample_backward(ptr(g), c, ptr(gi), c, b, h, w, ctx.size[0], ctx.size[1], c,
                                               1 if ctx.align else 0, stream()), "df_enc_upsample_backward")
            return gi, None, None
        gi = torch.ops.aten.upsample_bilinear2d_backward(g.contiguous(), list(ctx.size), list(ctx.in_shape), ctx.align,
                                                         None, None)
        return gi, None, None


def _upsample(x, size, align_corners: bool):
    """Bilinear resize.  CUDA fp32 tensors go through df_upsample_bilinear (every output element is a thread; torch's
    NCHW kernel loops over batch x channels inside one thread per output pixel: 64% of the first inference step and
    21% of a training step).  CPU tensors or other dtypes stay on torch."""
    if x.is_cuda and x.dtype == torch.float32:
        if torch.is_grad_enabled() and x.requires_grad:
            return _UpsampleFn.apply(x, size, align_corners)
        from .. import ops
        return ops.upsample_bilinear(x, size, align_corners)
    return F.interpolate(x, size=size, mode="bilinear", align_corners=align_corners)


class PSPModule(nn.Module):
    def __init__(self, features: int, out_features: int = 1024, sizes=(1, 2, 3, 6)):
        super().__init__()
        self.sizes = tuple(sizes)
        # stages.<i>.1.weight : index 0 is the (parameter-free) adaptive pooling
        self.stages = nn.ModuleList(
            nn.Sequential(nn.AdaptiveAvgPool2d((s, s)), nn.Conv2d(features, features, 1, bias=False)) for s in sizes)
        self.bottleneck = nn.Conv2d(features * (len(sizes) + 1), out_features, 1)

    def forward(self, feats):
        from . import ew
        if ew.usable(feats) and self.sizes == (1, 2, 3, 6):
            # training path on own kernels: the four adaptive pools in one pass, the concat written slice by slice (four resizes + one
            # pitched copy), ReLU as the epilogue of the bottleneck convolution
            pooled = ew.PyramidPoolFn.apply(feats)
            ys = [conv2d(stage[1], p) for stage, p in zip(self.stages, pooled)]
            return conv2d(self.bottleneck, ew.PyramidCatFn.apply(feats, *ys), act=1)
        hw = feats.shape[2:]
        pyramid = [_upsample(conv2d(stage[1], stage[0](feats)), hw, False) for stage in self.stages]
        return F.relu(conv2d(self.bottleneck, torch.cat(pyramid + [feats], 1)))


class PSPUpsample(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.conv = nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True),
                                  nn.Conv2d(cin, cout, 3, padding=1), nn.PReLU())

    def forward(self, x):
        # conv = [Upsample(x2, align_corners=True), Conv2d 3x3, PReLU]; indices kept for the checkpoint keys
        from . import ew
        x = _upsample(x, (x.shape[2] * 2, x.shape[3] * 2), True)
        y = conv2d(self.conv[1], x)
        return ew.PReLUFn.apply(y, self.conv[2].weight) if ew.usable(y) else self.conv[2](y)


class PSPNet(nn.Module):
    def __init__(self, n_classes=21, sizes=(1, 2, 3, 6), psp_size=512, deep_features_size=256, backend="resnet18"):
        super().__init__()
        self.feats = getattr(extractors, backend)()
        self.psp = PSPModule(psp_size, 1024, sizes)
        self.drop_1 = nn.Dropout2d(p=0.3)
        self.up_1 = PSPUpsample(1024, 256)
        self.up_2 = PSPUpsample(256, 64)
        self.up_3 = PSPUpsample(64, 64)
        self.drop_2 = nn.Dropout2d(p=0.15)
        self.final = nn.Sequential(nn.Conv2d(64, 32, 1), nn.LogSoftmax(dim=1))
        # unused by the pose path but part of the reference checkpoints
        self.classifier = nn.Sequential(nn.Linear(deep_features_size, 256), nn.ReLU(), nn.Linear(256, n_classes))

    def _drop(self, m: nn.Dropout2d, x):
        from . import ew
        if self.training and ew.usable(x):
            return ew.Dropout2dFn.apply(x, m.p)
        return m(x)

    def forward(self, x):
        from . import ew
        f, _ = self.feats(x)
        p = self._drop(self.drop_1, self.psp(f))
        p = self._drop(self.drop_2, self.up_1(p))
        p = self._drop(self.drop_2, self.up_2(p))
        y = conv2d(self.final[0], self.up_3(p))
        return ew.LogSoftmax32Fn.apply(y) if (ew.usable(y) and y.shape[1] == 32) else self.final[1](y)
