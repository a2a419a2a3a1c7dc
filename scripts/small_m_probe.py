"""Probe: does the tensor-core GEMM handle M < 256 (one partial CTA-pair tile), and how long does it take vs the fp32 kernel?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densefusion_b200 import ops
from densefusion_b200._C import check, lib, ptr, stream

torch.manual_seed(0)
dev = "cuda"
for (M, N, K) in ((64, 1024, 512), (96, 1024, 512), (128, 512, 1024), (200, 1024, 512), (32, 512, 128), (1, 1024, 512)):
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    bias = torch.randn(N, device=dev)
    want = torch.relu(A.double() @ W.double().T + bias.double())
    sw = ops.SplitWeight(W)
    for prec in ("3xtf32", "hybrid", "fp32"):
        guard = torch.full((M + 8, N), 7.0, device=dev)
        C = guard[:M]
        def run():
            if prec == "fp32":
                ops.gemm(A, W, bias, C, M=M, N=N, K=K, lda=K, ldw=K, ldc=N, relu=True)
            else:
                mode = ops.PRECISIONS[prec]
                hi, lo = sw.pairs() if mode == 3 else sw.split()
                check(lib.df_gemm_tc(ptr(A), K, ptr(hi), ptr(lo), K, ptr(bias), 0, ptr(C), N, M, N, K, 1, 0, 1, 0, 0, 0, None, mode, 0, stream()), "tc")
        run(); torch.cuda.synchronize()
        err = float((C.double() - want).abs().max() / want.abs().max())
        ok = float(guard[M:].min()) == 7.0 and float(guard[M:].max()) == 7.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        print(f"M={M} N={N} K={K} {prec}: err {err:.2e} guard {'ok' if ok else 'CLOBBERED'} {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
