"""Profiling driver (ncu): ONE tensor-core launch shape of the pose step, `reps` launches.

    python scripts/prof_case.py <case> [precision] [reps]
cases: tower1 tower2 tower3 conv5 conv5r conv6 pf2 up1 up2 bott l40 l41 l40b l41b l31 l1   (l40b / l41b: layer4 on the 15x15 maps of the
120 px bucket -- the balanced (tile, run) schedule); DF_PROF_CROPS = crops per head chunk (256)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densefusion_b200 import ops
from densefusion_b200.encoder import PackedEncoder, _pack_conv

dev = "cuda"
torch.manual_seed(0)
case = sys.argv[1]
mode = sys.argv[2] if len(sys.argv) > 2 else "hybrid16s"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
crops, n = int(os.environ.get("DF_PROF_CROPS", "256")), 500
rows = crops * n


def gemm(M, N, K, pooled=False, percrop=False, groups=1):
    A = torch.randn(M, K * groups, device=dev)
    W = ops.SplitWeight(torch.randn(groups * N, K, device=dev) / K ** 0.5)
    bias = torch.randn(crops if percrop else 1, N * groups, device=dev)
    C = torch.empty(M, N * groups, device=dev)
    part = torch.empty(M // n, 4, N, device=dev) if pooled else None
    for _ in range(reps):
        ops.gemm(A, W, bias, None if pooled else C, M=M, N=N, K=K, lda=K * groups, ldw=K, ldc=N * groups, relu=True, precision=mode,
                 bias_crop_stride=N * groups if percrop else 0, rows_per_crop=n, groups=groups, a_gs=K, w_gs=N * K, bias_gs=N, c_gs=N,
                 pool_partial=part)


def conv(B, H, W, Cin, Cout, dil, taps=9):
    x = torch.randn(B, H, W, Cin, device=dev)
    w = _pack_conv(torch.randn(Cout, Cin, 3 if taps == 9 else 1, 3 if taps == 9 else 1, device=dev) / (taps * Cin) ** 0.5)
    out = torch.empty(B, H, W, Cout, device=dev)
    for _ in range(reps):
        PackedEncoder._conv(x, w, out, taps=taps, dil=dil, act=1, mode=ops.PRECISIONS[mode])


{"tower1": lambda: gemm(rows, 1920, 384, percrop=True), "tower2": lambda: gemm(rows, 256, 640, groups=3),
 "tower3": lambda: gemm(rows, 128, 256, groups=3), "conv5": lambda: gemm(rows, 512, 256), "conv5r": lambda: gemm(rows, 512, 384),
 "conv6": lambda: gemm(rows, 1024, 512, pooled=True), "pf2": lambda: gemm(rows, 128, 64), "up1": lambda: gemm(25600, 2304, 1024),
 "up2": lambda: gemm(102400, 576, 256), "bott": lambda: conv(64, 20, 20, 512, 1024, 1, taps=1), "l40": lambda: conv(64, 20, 20, 512, 512, 1),
 "l41": lambda: conv(64, 20, 20, 512, 512, 4),
 "l40b": lambda: conv(96, 15, 15, 512, 512, 1), "l41b": lambda: conv(96, 15, 15, 512, 512, 4), "l31": lambda: conv(64, 20, 20, 256, 256, 2), "l1": lambda: conv(64, 40, 40, 64, 64, 1)}[case]()
torch.cuda.synchronize()
print("ok")
