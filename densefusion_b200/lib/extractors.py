"""BN-free dilated ResNet-18 trunk of the colour encoder.

Parameter names and shapes match the reference's lib/extractors.py:78-129 so reference checkpoints load unchanged; the
construction is table-driven rather than class-per-block.  Inference runs on densefusion_b200.encoder; this module graph is
the TRAINING path: its convolutions (with ReLU / skip-connection epilogues) go through lib/conv_tc.py, max pooling through
lib/ew.py -- torch ops only as the fallback for CPU tensors / the strict-fp32 parity mode."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .conv_tc import conv2d


class ResidualPair(nn.Module):
    """Two 3x3 convolutions with an identity or 1x1-projected skip (no normalisation layers)."""

    def __init__(self, cin: int, cout: int, stride: int, dilation: int, project: bool):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride=stride, padding=dilation, dilation=dilation, bias=False)
        self.conv2 = nn.Conv2d(cout, cout, 3, stride=1, padding=dilation, dilation=dilation, bias=False)
        self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False)) if project else None

    def forward(self, x):
        # conv2d(): tensor-core forward / data gradient in training (ReLU and the skip connection as epilogues of the convolution
        # kernel), the plain module + torch ops otherwise
        skip = x if self.downsample is None else conv2d(self.downsample[0], x)
        return conv2d(self.conv2, conv2d(self.conv1, x, act=1), act=1, residual=skip)


# (name, width, stride of the first pair, dilation of the later pairs); the first pair of a stage is
# never dilated (lib/extractors.py:99-112 does not forward `dilation` to it).
_STAGES = (("layer1", 64, 1, 1), ("layer2", 128, 2, 1), ("layer3", 256, 1, 2), ("layer4", 512, 1, 4))


class ResNet18Trunk(nn.Module):
    def __init__(self, pairs_per_stage=(2, 2, 2, 2)):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        width = 64
        for (name, cout, stride, dilation), count in zip(_STAGES, pairs_per_stage):
            blocks = [ResidualPair(width, cout, stride, 1, stride != 1 or width != cout)]
            blocks += [ResidualPair(cout, cout, 1, dilation, False) for _ in range(count - 1)]
            setattr(self, name, nn.Sequential(*blocks))
            width = cout
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                nn.init.normal_(m.weight, 0.0, math.sqrt(2.0 / fan))

    def forward(self, x):
        from . import conv_tc
        if conv_tc.ENABLED and x.is_cuda and torch.is_grad_enabled() and self.conv1.weight.requires_grad:
            x = x.contiguous(memory_format=torch.channels_last)     # NHWC storage for the tensor-core training convolutions
        from . import ew
        x = conv2d(self.conv1, x, act=1)
        x = ew.MaxPoolFn.apply(x) if ew.usable(x) else F.max_pool2d(x, 3, stride=2, padding=1)
        x = self.layer2(self.layer1(x))
        mid = self.layer3(x)
        return self.layer4(mid), mid


def resnet18(pretrained: bool = False) -> ResNet18Trunk:
    return ResNet18Trunk((2, 2, 2, 2))
