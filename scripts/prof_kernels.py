"""Run each hand-written hot kernel a few times at its bench shape (profiling driver for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densefusion_b200 import ops, synth
from densefusion_b200.encoder import PackedEncoder, _pack_conv

dev = "cuda"
torch.manual_seed(0)
crops, n = 128, 500
rows = crops * n
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3


def gemm_case(M, N, K, precision, pooled=False, percrop=False, groups=1):
    A = torch.randn(M, K * groups, device=dev)
    W = ops.SplitWeight(torch.randn(groups * N, K, device=dev) / K ** 0.5)
    W.split()
    bias = torch.randn(crops if percrop else 1, N * groups, device=dev)
    C = torch.empty(M, N * groups, device=dev)
    part = torch.empty(M // n, 4, N, device=dev) if pooled else None
    for _ in range(reps):
        ops.gemm(A, W, bias, None if pooled else C, M=M, N=N, K=K, lda=K * groups, ldw=K, ldc=N * groups, relu=True,
                 precision=precision, bias_crop_stride=N * groups if percrop else 0, rows_per_crop=n, groups=groups,
                 a_gs=K, w_gs=N * K, bias_gs=N, c_gs=N, pool_partial=part)


def conv_case(B, H, W, Cin, Cout, dil, mode):
    x = torch.randn(B, H, W, Cin, device=dev)
    w = _pack_conv(torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5)
    out = torch.empty(B, H, W, Cout, device=dev)
    for _ in range(reps):
        PackedEncoder._conv(x, w, out, taps=9, dil=dil, act=1, mode=mode)


# order == order of the cases in profiles/*_ncu_full_selected.csv
gemm_case(rows, 1920, 384, "hybrid", percrop=True)        # 1 tower layer 1 at the bench's chunk, bench-default arithmetic (the roofline kernel)
gemm_case(rows, 1920, 384, "3xtf32", percrop=True)        # 2 the same in 3xTF32
gemm_case(rows, 1024, 512, "hybrid", pooled=True)         # 3 conv6 + pool
conv_case(64, 40, 40, 1024, 256, 1, 3)                    # 4 up_1 convolution, 160x160 bucket (K = 9216, 16 accumulation runs), hybrid
conv_case(64, 20, 20, 512, 512, 4, 3)                     # 5 layer4.1 convolution, dilation 4, hybrid
conv_case(64, 80, 80, 256, 64, 1, 1)                      # 6 up_2 convolution (64 output channels: stays on 3xTF32)
# loss (ADD-S) and kNN at config C1 shapes (32 crops)
g = torch.Generator().manual_seed(1)
B = 32
pr = torch.randn(B, n, 4, generator=g).to(dev); pt = (torch.randn(B, n, 3, generator=g) * 0.02).to(dev)
pc = (torch.rand(B, n, 1, generator=g) * 0.9 + 0.05).to(dev)
model = (torch.randn(B, 500, 3, generator=g) * 0.05).to(dev); target = model + torch.tensor([0., 0., 0.8], device=dev)
cloud = (torch.randn(B, n, 3, generator=g) * 0.05 + torch.tensor([0., 0., 0.8])).to(dev)
obj = torch.full((B,), 12, dtype=torch.int64, device=dev)
for _ in range(reps):
    ops.loss_forward(pr, pt, pc, target, model, cloud, cloud, obj, ops.sym_mask(synth.YCB_SYM), True, 0.015)
ref = target[0].t().contiguous()[None]
qry = (torch.randn(1, 3, 2_000_000, generator=g) * 0.05 + torch.tensor([0., 0., 0.8]).view(1, 3, 1)).to(dev)
for _ in range(reps):
    ops.knn(ref, qry, 1)
torch.cuda.synchronize()
print("ok")
