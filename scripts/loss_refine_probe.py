"""Latency of the refiner loss (K3 at P = 1, lib/loss_refiner.py:12-62) with the hypothesis shared by a cluster of 8 CTAs (default)
or one CTA per crop (DF_LOSS_CLUSTER=0): Q = R = M model / target points, symmetric object (1-NN scan), B crops."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densefusion_b200 import ops

dev = "cuda"
out = {"cluster": os.environ.get("DF_LOSS_CLUSTER", "1")}
for (B, M) in ((1, 500), (1, 2600), (8, 500), (256, 500)):
    g = torch.Generator(device=dev).manual_seed(B * 7 + M)
    r = torch.randn(B, 1, 4, device=dev, generator=g)
    t = torch.randn(B, 1, 3, device=dev, generator=g) * 0.01
    tgt = torch.randn(B, M, 3, device=dev, generator=g) * 0.1
    mdl = torch.randn(B, M, 3, device=dev, generator=g) * 0.1
    pts = torch.randn(B, 500, 3, device=dev, generator=g) * 0.1
    idx = torch.zeros(B, dtype=torch.int64, device=dev)

    def run():
        return ops.loss_forward(r, t, None, tgt, mdl, None, pts, idx, 1, True, 0.015)
    for _ in range(5):
        st = run()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(20):
            run()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    out[f"B={B} M={M}"] = {"us_per_call": round(e0.elapsed_time(e1) / 200 * 1000, 2), "dis0": float(st.dis_sel[0])}
print(json.dumps(out))
