import sys, os, json
sys.path.insert(0, "/root/repo")
import torch
from densefusion_b200 import ops
from densefusion_b200.encoder import PackedEncoder, _pack_conv
def timeit(fn, reps=20):
    for _ in range(3): fn(0); fn(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i & 1)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out={"dbg": os.environ.get("DF_TC_DBG","0")}
for (B,H,W,ci,co,d) in ((64,40,40,64,64,1),(64,20,20,128,128,1),(64,20,20,256,256,2)):
    x=[torch.randn(B,H,W,ci,device="cuda") for _ in range(2)]
    w=_pack_conv(torch.randn(co,ci,3,3,device="cuda")/(9*ci)**0.5)
    o=[torch.empty(B,H,W,co,device="cuda") for _ in range(2)]
    def run(i): PackedEncoder._conv(x[i], w, o[i], taps=9, dil=d, act=1, mode=6)
    out[f"{B}x{H}x{W} {ci}->{co} d{d}"]=round(timeit(run),4)
print(json.dumps(out), flush=True)
