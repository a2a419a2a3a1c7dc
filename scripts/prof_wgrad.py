"""Profiling driver for ncu: the tensor-core weight gradient (df_conv_wgrad_tc) at two training shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densefusion_b200._C import check, lib, ptr, stream

CASES = ["wgrad transpose dY: layer4.1 3x3 dil 4 512->512, 6 crops of 10x10", "wgrad transpose X (3 shifted planes, hi/lo)",
         "wgrad GEMM (split-K 3xTF32): M=512 N=4608 K=pixels", "wgrad transpose dY: head tower layer 1, 8000 rows x 1920",
         "wgrad transpose X: 8000 rows x 384", "wgrad GEMM (split-K 3xTF32): M=1920 N=384 K=8000"]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for (B, H, W, Cin, Cout, taps, dil) in ((6, 10, 10, 512, 512, 9, 4), (1, 1, 8000, 384, 1920, 1, 1)):
    x, dy = torch.randn(B, H, W, Cin, device="cuda"), torch.randn(B, H, W, Cout, device="cuda")
    scratch = torch.empty(int(lib.df_conv_wgrad_scratch_floats(B, H, W, Cin, Cout, taps, dil)), device="cuda")
    out = torch.empty(Cout, taps * Cin, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for r in range(reps):
        if r == reps - 1:
            e0.record()
        check(lib.df_conv_wgrad_tc(ptr(x), Cin, ptr(dy), Cout, B, H, W, Cin, Cout, taps, dil, ptr(scratch), ptr(out), stream()), "wgrad")
    e1.record()
    torch.cuda.synchronize()
    fl = 2.0 * B * H * W * Cin * Cout * taps
    print(f"{(B, H, W, Cin, Cout, taps, dil)}: {e0.elapsed_time(e1) * 1e3:.1f} us  {fl / e0.elapsed_time(e1) / 1e9:.1f} TFLOP/s")
print("ok")
if len(sys.argv) > 2:
    print("\n".join(CASES))
