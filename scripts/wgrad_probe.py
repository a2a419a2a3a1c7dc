"""Probe df_conv_wgrad_tc case by case, each in its own process (a trap poisons the CUDA context)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [(3, 10, 10, 256, 512, 9, 4), (2, 15, 15, 128, 64, 9, 2), (2, 15, 15, 128, 128, 9, 2), (2, 15, 15, 256, 64, 9, 2),
         (2, 16, 16, 128, 64, 9, 2), (2, 15, 15, 128, 256, 9, 2), (2, 20, 12, 64, 64, 9, 1), (4, 10, 10, 512, 1024, 1, 1),
         (1, 1, 2500, 384, 1920, 1, 1), (1, 1, 999, 64, 128, 1, 1), (2, 15, 15, 128, 64, 1, 1)]

if len(sys.argv) > 1:
    import torch
    import torch.nn.functional as F
    from densefusion_b200._C import check, lib, ptr, stream
    B, H, W, Cin, Cout, taps, dil = [int(v) for v in sys.argv[1:8]]
    g = torch.Generator().manual_seed(1)
    k = 3 if taps == 9 else 1
    x, dy = torch.randn(B, Cin, H, W, generator=g), torch.randn(B, Cout, H, W, generator=g)
    w = torch.zeros(Cout, Cin, k, k, dtype=torch.float64, requires_grad=True)
    (F.conv2d(x.double(), w, padding=dil * (k // 2), dilation=dil) * dy.double()).sum().backward()
    xn, dyn = x.permute(0, 2, 3, 1).contiguous().cuda(), dy.permute(0, 2, 3, 1).contiguous().cuda()
    scratch = torch.empty(int(lib.df_conv_wgrad_scratch_floats(B, H, W, Cin, Cout, taps, dil)), device="cuda")
    out = torch.empty(Cout, taps * Cin, device="cuda")
    check(lib.df_conv_wgrad_tc(ptr(xn), Cin, ptr(dyn), Cout, B, H, W, Cin, Cout, taps, dil, ptr(scratch), ptr(out), stream()), "wgrad")
    torch.cuda.synchronize()
    got = out.view(Cout, taps, Cin).permute(0, 2, 1).reshape(Cout, Cin, k, k).cpu().double()
    print("err %.3e" % float((got - w.grad).abs().max() / w.grad.abs().max()))
else:
    for c in CASES:
        r = subprocess.run([sys.executable, __file__] + [str(v) for v in c], capture_output=True, text=True)
        tail = (r.stdout.strip().splitlines() or [""])[-1] if r.returncode == 0 else (r.stderr.strip().splitlines() or ["?"])[-1][:150]
        print(c, "rc", r.returncode, tail, flush=True)
