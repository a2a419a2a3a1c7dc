"""Profiling driver (ncu): the tensor-core GEMM in the "hybrid16s" arithmetic at the bench's tower-1 shape and as the layer4.0 convolution."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densefusion_b200 import ops
from densefusion_b200.encoder import PackedEncoder, _pack_conv

dev = "cuda"
torch.manual_seed(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "hybrid16s"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
crops, n = 128, 500
rows = crops * n
A = torch.randn(rows, 384, device=dev)
W = ops.SplitWeight(torch.randn(1920, 384, device=dev) / 384 ** 0.5)
bias = torch.randn(crops, 1920, device=dev)
C = torch.empty(rows, 1920, device=dev)
for _ in range(reps):
    ops.gemm(A, W, bias, C, M=rows, N=1920, K=384, lda=384, ldw=384, ldc=1920, relu=True, precision=mode, bias_crop_stride=1920, rows_per_crop=n)
x = torch.randn(64, 20, 20, 512, device=dev)
w = _pack_conv(torch.randn(512, 512, 3, 3, device=dev) / (9 * 512) ** 0.5)
out = torch.empty(64, 20, 20, 512, device=dev)
for _ in range(reps):
    PackedEncoder._conv(x, w, out, taps=9, dil=1, act=1, mode=ops.PRECISIONS[mode])
torch.cuda.synchronize()
print("ok")
