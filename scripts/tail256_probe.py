"""256-wide tiles over an N that is not a multiple of 256 (DF_TC_TAIL256, gemm_tc.cu::pair_tile_width): the half-masked tail tile must
leave everything past column N untouched and match float64 like the other widths.  One JSON line per case.

    DF_TC_TAIL256=2 python scripts/tail256_probe.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densefusion_b200 import ops

torch.manual_seed(1)
for (M, N, K, ldc, relu, percrop) in [(8000, 1920, 384, 1920, True, True), (3001, 1920, 384, 2048, True, False), (2500, 1408, 128, 1536, False, False),
                                      (700, 1152, 1024, 1152, True, False)]:
    n = 500
    A = torch.randn(M, K, device="cuda")
    W = ops.SplitWeight(torch.randn(N, K, device="cuda") / K ** 0.5)
    crops = (M + n - 1) // n
    bias = torch.randn(crops if percrop else 1, N, device="cuda")
    C = torch.full((M, ldc), 777.0, device="cuda")
    ops.gemm(A, W, bias, C, M=M, N=N, K=K, lda=K, ldw=K, ldc=ldc, relu=relu, precision="hybrid16s",
             bias_crop_stride=N if percrop else 0, rows_per_crop=n)
    torch.cuda.synchronize()
    ref = A.double() @ W.w.double().t() + (bias.double().repeat_interleave(n, 0)[:M] if percrop else bias.double())
    if relu:
        ref = torch.relu(ref)
    err = float((C[:, :N].double() - ref).abs().max() / ref.abs().max())
    untouched = bool((C[:, N:] == 777.0).all().item()) if ldc > N else True
    print(json.dumps({"M": M, "N": N, "K": K, "ldc": ldc, "tail256": os.environ.get("DF_TC_TAIL256", ""), "tstore": os.environ.get("DF_TC_TSTORE", ""),
                      "err_vs_f64": err, "columns_past_N_untouched": untouched, "ok": err < 5e-6 and untouched}), flush=True)
