"""Balanced schedule of the long-K convolutions (gemm_tc.cu, QSched / FORM 4): result against a float64 convolution, run-to-run
bit equality and time per launch, for the layer3 / layer4 shapes of the bench step.  DF_TC_SPLIT=0 in the environment gives the
round-robin schedule for comparison (scripts/gpu_split_ab.sh runs both)."""
import ctypes
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from densefusion_b200 import _C, ops  # noqa: E402
from densefusion_b200.encoder import PackedEncoder, _pack_conv  # noqa: E402

dev = "cuda"
CASES = [(96, 15, 15, 512, 512, 1), (96, 15, 15, 512, 512, 4), (64, 20, 20, 512, 512, 1), (64, 20, 20, 512, 512, 4),
         (96, 10, 10, 512, 512, 1), (64, 20, 20, 256, 256, 2), (96, 15, 15, 256, 256, 2), (96, 15, 15, 256, 512, 1),
         (64, 20, 20, 256, 512, 1), (40, 20, 20, 512, 512, 1), (37, 15, 15, 512, 512, 2)]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for (B, H, W, Cin, Cout, dil) in CASES:
    g = torch.Generator(device=dev).manual_seed(B * 1000 + H + Cin + Cout + dil)
    x = torch.randn(B, H, W, Cin, device=dev, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device=dev, generator=g) / (9 * Cin) ** 0.5
    bias = torch.randn(Cout, device=dev, generator=g)
    res = torch.randn(B, H, W, Cout, device=dev, generator=g)
    pw = _pack_conv(w)
    code = ops.PRECISIONS["hybrid16s"]
    outs = []
    for rep in range(3):
        o = torch.full((B, H, W, Cout), 7.0, device=dev)
        PackedEncoder._conv(x, pw, o, taps=9, dil=dil, bias=bias, residual=res, act=1, mode=code)
        torch.cuda.synchronize()
        outs.append(o)
    want = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), bias.double(), padding=dil, dilation=dil).permute(0, 2, 3, 1) + res.double()
    want = torch.relu(want)
    err = float((outs[0].double() - want).abs().max() / want.abs().max())
    o2 = torch.empty(B, H, W, Cout, device=dev)
    ms = timeit(lambda: PackedEncoder._conv(x, pw, o2, taps=9, dil=dil, bias=bias, residual=res, act=1, mode=code))
    info = (ctypes.c_int * (5 + 4 * 74))()
    rc = _C.lib.df_conv_tc_schedule(B, H, W, Cin, Cout, dil, 74, ctypes.cast(info, ctypes.c_void_p))
    print(json.dumps({"case": f"{B}x{H}x{W} {Cin}->{Cout} dil {dil}", "split_env": os.environ.get("DF_TC_SPLIT", "1"), "planned": rc,
                      "rr_max_kb": info[0], "split_max_kb": info[1], "ms": round(ms, 4), "err_vs_f64": err,
                      "bit_equal_runs": bool(torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]))}), flush=True)
