"""A/B timing of the tensor-core GEMM / convolution at the bench shapes: one precision mode against another, CUDA events,
each launch on fresh operands (two operand sets larger than L2 alternate).  Prints one JSON line per case.

    python scripts/gemm_ab.py [modeA modeB ...]        default: hybrid16 hybrid16s
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densefusion_b200 import ops
from densefusion_b200.encoder import PackedEncoder, _pack_conv

dev = "cuda"
modes = sys.argv[1:] or ["hybrid16", "hybrid16s"]
torch.manual_seed(0)
REPS = 20
ONLY = [t for t in os.environ.get("DF_AB_ONLY", "").split(",") if t]      # substrings of the case names to run (default: all)


def timeit(fn):
    for _ in range(3):
        fn(0); fn(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(REPS):
        fn(i & 1)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REPS


def gemm_case(name, M, N, K, pooled=False, percrop=False, groups=1, n=500):
    if ONLY and not any(t in name for t in ONLY):
        return
    A = [torch.randn(M, K * groups, device=dev) for _ in range(2)]
    W = ops.SplitWeight(torch.randn(groups * N, K, device=dev) / K ** 0.5)
    crops = M // n
    bias = torch.randn(crops if percrop else 1, N * groups, device=dev)
    C = [torch.empty(M, N * groups, device=dev) for _ in range(2)]
    part = torch.empty(max(crops, 1), 4, N, device=dev) if pooled else None
    out = {}
    res = {}
    for mode in modes:
        def run(i, mode=mode):
            ops.gemm(A[i], W, bias, None if pooled else C[i], M=M, N=N, K=K, lda=K * groups, ldw=K, ldc=N * groups, relu=True,
                     precision=mode, bias_crop_stride=N * groups if percrop else 0, rows_per_crop=n, groups=groups,
                     a_gs=K, w_gs=N * K, bias_gs=N, c_gs=N, pool_partial=part)
        ms = timeit(run)
        run(0)
        torch.cuda.synchronize()
        res[mode] = (part if pooled else C[0]).clone()
        out[mode] = {"ms": round(ms, 4), "algorithmic_tflops": round(2.0 * M * N * K * groups / ms * 1e-9, 1)}
    eq = {m: bool(torch.equal(res[m], res[modes[0]])) for m in modes[1:]}
    if not pooled and M * N * groups <= 130_000_000:            # error against float64 (cuBLAS DGEMM on the device)
        ref = torch.relu(torch.einsum("mgk,gnk->mgn", A[0].double().view(M, groups, K), W.w.double().view(groups, N, K)).reshape(M, -1)
                         + (bias.double().repeat_interleave(n, 0)[:M] if percrop else bias.double()))
        for mode in modes:
            out[mode]["err_vs_f64"] = float((res[mode].double() - ref).abs().max() / ref.abs().max())
        del ref
    print(json.dumps({"case": name, "shape": f"M={M} N={N} K={K} groups={groups}", **out, "bit_equal_to_first": eq}), flush=True)


def conv_case(name, B, H, W, Cin, Cout, dil):
    if ONLY and not any(t in name for t in ONLY):
        return
    x = [torch.randn(B, H, W, Cin, device=dev) for _ in range(2)]
    w = _pack_conv(torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5)
    o = [torch.empty(B, H, W, Cout, device=dev) for _ in range(2)]
    out, res = {}, {}
    for mode in modes:
        code = ops.PRECISIONS[mode]
        def run(i, code=code):
            PackedEncoder._conv(x[i], w, o[i], taps=9, dil=dil, act=1, mode=code)
        ms = timeit(run)
        run(0)
        torch.cuda.synchronize()
        res[mode] = o[0].clone()
        out[mode] = {"ms": round(ms, 4), "dense_tflops": round(2.0 * B * H * W * Cout * Cin * 9 / ms * 1e-9, 1)}
    eq = {m: bool(torch.equal(res[m], res[modes[0]])) for m in modes[1:]}
    print(json.dumps({"case": name, "shape": f"{B}x{H}x{W} {Cin}->{Cout} dil {dil}", **out, "bit_equal_to_first": eq}), flush=True)


rows = 128 * 500
gemm_case("tower1 (bench roofline kernel)", rows, 1920, 384, percrop=True)
gemm_case("tower1 at 256 crops (bench chunk)", 2 * rows, 1920, 384, percrop=True)
gemm_case("tower2 grouped", rows, 256, 640, groups=3)
gemm_case("tower3 grouped", rows, 128, 256, groups=3)
gemm_case("conv5", rows, 512, 256)
gemm_case("conv5 K=384", rows, 512, 384)
gemm_case("conv6 pooled", rows, 1024, 512, pooled=True)
gemm_case("pf conv2 (K=64)", rows, 128, 64)
gemm_case("up_1 low-res GEMM", 25600, 2304, 1024)
gemm_case("up_2 low-res GEMM", 102400, 576, 256)
gemm_case("bottleneck K=512", 25600, 1024, 512)
conv_case("layer4.1 dil 4", 64, 20, 20, 512, 512, 4)
conv_case("layer4.0 dil 1", 64, 20, 20, 512, 512, 1)
conv_case("layer3.1 dil 2", 64, 20, 20, 256, 256, 2)
conv_case("layer2.1", 64, 20, 20, 128, 128, 1)
conv_case("layer4.1 dil 4, 80 px crops", 96, 10, 10, 512, 512, 4)
