"""Accuracy of the "hybrid16s" arithmetic (two fp16 planes per operand, power-of-two scales) against float64 over operand scales,
next to "hybrid16" and the exact-fp32 kernel; plus fixed activation scales (a_log2) to show the range the sampled scale protects.

    python scripts/h16s_probe.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from densefusion_b200 import ops

dev = "cuda"
g = torch.Generator().manual_seed(3)
M, N, K = 2048, 512, 1024


def run(x, w, mode, a_log2=None, relu_in=False):
    C = torch.empty(M, N, device=dev)
    ops.gemm(x, ops.SplitWeight(w), None, C, M=M, N=N, K=K, lda=K, ldw=K, ldc=N, relu=False, precision=mode, a_log2=a_log2)
    return C


for relu_in in (False, True):
    for sa, sw in ((1.0, 1.0), (1e-2, 1.0), (1e-4, 1.0), (1e-7, 1.0), (1e4, 1.0), (1e6, 1.0), (1.0, 1e-3), (1.0, 1e3), (1e-6, 1e5), (1e5, 1e-6)):
        x = torch.randn(M, K, generator=g) * sa
        if relu_in:
            x = torch.relu(x)
        w = torch.randn(N, K, generator=g) / K ** 0.5 * sw
        x, w = x.to(dev), w.to(dev)
        ref = x.double() @ w.double().t()
        d = float(ref.abs().max())
        row = {"relu_input": relu_in, "scale_a": sa, "scale_w": sw}
        for mode in ("fp32", "hybrid16", "hybrid16s"):
            row[mode] = round(float((run(x, w, mode).double() - ref).abs().max()) / d, 10)
        for k in (0, 5):
            row[f"hybrid16s_fixed_2^{k}"] = round(float((run(x, w, "hybrid16s", a_log2=k).double() - ref).abs().max()) / d, 10)
        print(json.dumps(row), flush=True)
# one huge outlier in an otherwise O(1) operand (the sample cannot see it): graceful up to 2^16 / 2^5 x the sampled maximum
x = torch.randn(M, K, generator=g)
w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
for big in (1e2, 1e3, 4e3, 1e4):
    x2 = x.clone()
    x2[1234, 77] = big
    x2 = x2.to(dev)
    ref = x2.double() @ w.double().t()
    err = (run(x2, w, "hybrid16s").double() - ref).abs().max() / ref.abs().max()
    print(json.dumps({"outlier": big, "hybrid16s": float(err)}), flush=True)
