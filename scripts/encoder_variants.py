"""Time the colour encoder (torch/cuDNN) in its layout / precision variants on the bench crop mix."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densefusion_b200 import synth
from densefusion_b200.lib.network import PoseNet

dev = "cuda"
net = PoseNet(500, 21)
net.load_state_dict(synth.synth_state_dict(synth.shapes_of(net), 0))
cnn = net.cnn.eval().requires_grad_(False).to(dev)
out = {}
base = None
for name, cl, tf32, dtype in [("nchw_fp32", False, False, torch.float32), ("nhwc_fp32", True, False, torch.float32),
                               ("nchw_tf32", False, True, torch.float32), ("nhwc_tf32", True, True, torch.float32),
                               ("nhwc_bf16", True, True, torch.bfloat16)]:
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    m = cnn.to(memory_format=torch.channels_last if cl else torch.contiguous_format).to(dtype)
    res = {}
    for hw, b in ((80, 96), (120, 96), (160, 64)):
        x = torch.randn(b, 3, hw, hw, device=dev, dtype=dtype)
        if cl:
            x = x.contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            for _ in range(3):
                y = m(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                y = m(x)
            e1.record()
            torch.cuda.synchronize()
        res[f"{hw}x{hw}x{b}"] = e0.elapsed_time(e1) / 5
        if hw == 80:
            yf = y.float()
            if base is None:
                base = yf.clone()
            res["max_abs_diff_vs_nchw_fp32"] = float((yf - base).abs().max())
    res["total_ms_256_crops"] = sum(v for k, v in res.items() if "x" in k and not k.startswith("max"))
    out[name] = res
    print(name, json.dumps(res), flush=True)
json.dump(out, open("gpurun_out/encoder_variants.json", "w"), indent=1)
