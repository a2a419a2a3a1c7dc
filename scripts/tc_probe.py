"""Timing probes for the persistent tcgen05 GEMM: separate MMA-bound / store-bound / staging-bound regimes."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densefusion_b200 import ops

dev = "cuda"
crops, n = 32, 500
rows = crops * n


def t_ms(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def case(name, M, N, K, precision, pooled=False, variant=0):
    A = torch.randn(M, K, device=dev)
    W = ops.SplitWeight(torch.randn(N, K, device=dev) / K ** 0.5)
    W.split()
    bias = torch.randn(N, device=dev)
    C = torch.empty(M, N, device=dev)
    part = torch.empty(M // n, 4, N, device=dev) if pooled else None

    def run():
        ops.TC_VARIANT = variant
        ops.gemm(A, W, bias, None if pooled else C, M=M, N=N, K=K, lda=K, ldw=K, ldc=N, relu=True, precision=precision,
                 rows_per_crop=n, pool_partial=part)
        ops.TC_VARIANT = 0
    ms = t_ms(run)
    passes = 3 if precision == "3xtf32" else 1
    tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
    print(f"{name:34s} M={M} N={N} K={K} {precision:7s} pooled={int(pooled)} v={variant}: {ms*1e3:8.1f} us  alg {tf:7.1f} TF/s  executed {tf*passes:7.1f} TF/s", flush=True)


variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4, 5, 6]
for v in variants:
    case("tower1 store", rows, 1920, 384, "3xtf32", variant=v)
    case("tower1 tf32 store", rows, 1920, 384, "tf32", variant=v)
    case("conv6 pooled", rows, 1024, 512, "3xtf32", pooled=True, variant=v)
    case("conv5 store", rows, 512, 256, "3xtf32", variant=v)
    case("bigK store", rows, 1920, 1536, "3xtf32", variant=v)
    case("bigK tf32 store", rows, 1920, 1536, "tf32", variant=v)
    case("large M tower1 store", rows * 4, 1920, 384, "3xtf32", variant=v)
