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


def finish_case(B, h, w, C):
    from densefusion_b200._C import check, lib, ptr, stream
    z = torch.randn(B, h, w, 9 * C, device=dev)
    bias, slope = torch.randn(C, device=dev), torch.full((1,), 0.25, device=dev)
    out = torch.empty(B, 2 * h, 2 * w, C, device=dev)
    for _ in range(reps):
        check(lib.df_enc_upconv_finish(ptr(z), 9 * C, ptr(bias), ptr(slope), ptr(out), C, B, h, w, C, stream()), "upconv_finish")


# order == order of the cases in profiles/*_ncu_full_selected.csv (labels: CASES)
CASES = ["tower1 M=64000 N=1920 K=384 per-crop bias, hybrid16 (bench roofline kernel)", "the same, hybrid", "the same, 3xtf32",
         "conv6 M=64000 N=1024 K=512 pooled, hybrid16", "up_1 at the low resolution: GEMM M=25600 N=2304 K=1024, hybrid16",
         "layer4.1 conv 3x3 dil 4 512->512 on 64x20x20 (tap skipping), hybrid16", "layer4.0 conv 3x3 dil 1 512->512 on 64x20x20, hybrid16",
         "layer1 conv 3x3 64->64 on 64x40x40 (64 output channels: 3xTF32)", "up_1 finish: 9 shifted bilinear samples, 64x20x20 -> 64x40x40x256",
         "ADD-S loss, 32 crops", "kNN R=500 Q=2e6"]
gemm_case(rows, 1920, 384, "hybrid16", percrop=True)
gemm_case(rows, 1920, 384, "hybrid", percrop=True)
gemm_case(rows, 1920, 384, "3xtf32", percrop=True)
gemm_case(rows, 1024, 512, "hybrid16", pooled=True)
gemm_case(25600, 2304, 1024, "hybrid16")
conv_case(64, 20, 20, 512, 512, 4, 4)
conv_case(64, 20, 20, 512, 512, 1, 4)
conv_case(64, 40, 40, 64, 64, 1, 1)
finish_case(64, 20, 20, 256)
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
if len(sys.argv) > 2:
    print("\n".join(CASES))
