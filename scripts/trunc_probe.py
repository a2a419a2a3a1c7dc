"""How does the tensor core round when it adds products into the fp32 TMEM accumulator?  One long accumulation chain
(DF_TC_RUN_STEPS=1000000: no chunking), signed relative error against float64 for all-positive, all-negative and mixed sums."""
import json, os, sys
os.environ.setdefault("DF_TC_RUN_STEPS", "1000000")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densefusion_b200 import ops

g = torch.Generator().manual_seed(0)
M, N = 1024, 256
for K in (512, 4096):
    a = torch.rand(M, K, generator=g) + 0.5
    w = (torch.rand(N, K, generator=g) + 0.5) / K
    am = torch.randn(M, K, generator=g)
    wm = torch.randn(N, K, generator=g) / K ** 0.5
    for prec in ("hybrid16s", "hybrid16", "3xtf32", "hybrid"):
        row = {"K": K, "precision": prec}
        for name, (x, y) in {"pos": (a, w), "neg": (-a, w), "mixed": (am, wm), "relu_mixed": (am.clamp(min=0), wm)}.items():
            got = ops.linear(x.cuda(), y.cuda(), None, precision=prec).double().cpu()
            want = x.double() @ y.double().t()
            e = got - want
            scale = want.abs().mean()
            row[name] = {"mean_signed_rel": float((e / want.abs().clamp(min=1e-30) * torch.sign(want)).mean()) if name in ("pos", "neg") else None,
                         "mean_err_over_scale": float(e.mean() / scale), "mean_err_times_sign_over_scale": float((e * torch.sign(want)).mean() / scale),
                         "rms_over_scale": float(e.pow(2).mean().sqrt() / scale)}
        print(json.dumps(row), flush=True)
