"""conv1 (7x7/2): df_enc_conv1_tc (patches gathered by the GEMM kernel's stagers) against df_enc_im2col_conv1 + df_gemm_tc, ms per call."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densefusion_b200 import ops
from densefusion_b200._C import check, lib, ptr, stream

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for (B, H, W) in ((64, 160, 160), (96, 120, 120), (96, 80, 80)):
    img = torch.randn(B, 3, H, W, device="cuda")
    w1 = torch.zeros(64, 160, device="cuda"); w1[:, :147] = torch.randn(64, 147, device="cuda") * 0.05
    sw = ops.SplitWeight(w1); planes, scale = sw.planes16s()
    Ho, Wo = H // 2, W // 2
    y = torch.empty(B, Ho, Wo, 64, device="cuda"); a0 = torch.empty(B * Ho * Wo, 160, device="cuda")
    t_g = timeit(lambda: check(lib.df_enc_conv1_tc(ptr(img), B, H, W, ptr(planes), ptr(scale), ptr(y), 64, 64, 1, stream()), "c1"))
    t_i = timeit(lambda: check(lib.df_enc_im2col_conv1(ptr(img), ptr(a0), B, H, W, 160, stream()), "im2col"))
    t_m = timeit(lambda: ops.gemm(a0, sw, None, y.view(-1, 64), M=B * Ho * Wo, N=64, K=160, lda=160, ldw=160, ldc=64, relu=True, precision="hybrid16s"))
    print(json.dumps({"shape": [B, H, W], "gathered_ms": round(t_g, 4), "im2col_ms": round(t_i, 4), "gemm_ms": round(t_m, 4)}), flush=True)
