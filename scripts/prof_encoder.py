"""Per-layer device time of the tensor-core encoder (CUDA events around every conv / GEMM / helper launch)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from densefusion_b200 import ops, encoder as E
from densefusion_b200._C import lib

dev = torch.device("cuda", 0)
precision = sys.argv[1] if len(sys.argv) > 1 else "hybrid16"
est, ref, _, _ = bench.build_modules(dev)
enc = E.PackedEncoder(est.cnn)
records = []


def timed(label, fn, *a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(*a, **k); e1.record()
    records.append((label, e0, e1))
    return r


orig_conv, orig_gemm = E.PackedEncoder._conv, ops.gemm


def conv(x, w, out, **k):
    cout = k.get("cout") or out.shape[3]
    flops = 2.0 * x.shape[0] * x.shape[1] * x.shape[2] * k["taps"] * x.shape[3] * cout
    return timed(("conv", tuple(x.shape), cout, k["taps"], k.get("dil", 1), flops), orig_conv, x, w, out, **k)


def gemm(A, W, b, C, **k):
    return timed(("gemm", k["M"], k["N"], k["K"], k.get("precision", "fp32"), 2.0 * k["M"] * k["N"] * k["K"]), orig_gemm, A, W, b, C, **k)


E.PackedEncoder._conv = staticmethod(conv)
ops.gemm = gemm
for name in ("df_enc_im2col_conv1", "df_enc_maxpool", "df_enc_im2col_s2", "df_enc_pyramid_pool", "df_enc_pyramid_sum", "df_enc_upconv_finish", "df_enc_upsample", "df_enc_log_softmax32"):
    f = getattr(lib, name)
    setattr(lib, name, (lambda f, name: lambda *a: timed((name,), f, *a))(f, name))

for b, hw in ((96, 80), (96, 120), (64, 160)):
    img = torch.randn(b, 3, hw, hw, device=dev)
    for _ in range(2):
        enc.forward(img, precision)
    torch.cuda.synchronize()
    records.clear()
    enc.forward(img, precision)
    torch.cuda.synchronize()
    tot = sum(e0.elapsed_time(e1) for _, e0, e1 in records)
    print(f"\n# bucket {b} x {hw}x{hw}: {tot:.3f} ms over {len(records)} launches")
    for lab, e0, e1 in records:
        ms = e0.elapsed_time(e1)
        tf = f"{lab[-1] / ms / 1e9:7.1f} TF/s" if lab[0] in ("conv", "gemm") else ""
        print(f"{ms*1e3:9.1f} us {100*ms/tot:5.1f}%  {tf}  {lab[:-1] if lab[0] in ('conv','gemm') else lab}")
