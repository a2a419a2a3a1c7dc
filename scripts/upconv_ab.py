"""Times df_enc_upconv_finish at the bench shapes (run once with DF_UPCONV_SMEM=0 and once with =1) and prints a checksum."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densefusion_b200._C import check, lib, ptr, stream

dev = "cuda"
torch.manual_seed(0)
out_all = {}
for (B, h, w, C) in ((64, 20, 20, 256), (96, 15, 15, 256), (96, 10, 10, 256), (64, 40, 40, 64), (96, 30, 30, 64), (96, 20, 20, 64), (3, 7, 9, 64)):
    z = [torch.randn(B, h, w, 9 * C, device=dev) for _ in range(2)]
    bias, slope = torch.randn(C, device=dev), torch.full((1,), 0.25, device=dev)
    out = torch.empty(B, 2 * h, 2 * w, C, device=dev)

    def run(i):
        check(lib.df_enc_upconv_finish(ptr(z[i]), 9 * C, ptr(bias), ptr(slope), ptr(out), C, B, h, w, C, stream()), "upconv_finish")
    for i in range(4):
        run(i & 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        run(i & 1)
    e1.record()
    torch.cuda.synchronize()
    run(0)
    torch.cuda.synchronize()
    out_all[f"{B}x{h}x{w}x{C}"] = {"ms": round(e0.elapsed_time(e1) / 20, 4), "checksum": float(out.double().sum()), "absmax": float(out.abs().max()),
                                  "hash": int(torch.sum(out.view(torch.int32).long() % 1000003))}
print(json.dumps({"smem": os.environ.get("DF_UPCONV_SMEM", "1"), **out_all}))
