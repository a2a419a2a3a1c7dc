import os, sys, torch
sys.path.insert(0, "/root/repo")
from densefusion_b200.encoder import PackedEncoder, _pack_conv
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/it
for (B,H,W,Ci,Co,d) in [(64,40,40,1024,256,1),(64,80,80,256,64,1),(64,20,20,512,512,1),(96,30,30,64,64,1),(64,20,20,256,256,2)]:
    x=torch.randn(B,H,W,Ci,device="cuda"); w=_pack_conv(torch.randn(Co,Ci,3,3,device="cuda")/(9*Ci)**.5); o=torch.empty(B,H,W,Co,device="cuda")
    for mode,name in ((3,"hybrid"),(1,"3xtf32")):
        ms=t(lambda: PackedEncoder._conv(x,w,o,taps=9,dil=d,act=1,mode=mode))
        print(f"{(B,H,W,Ci,Co,d)} {name}: {ms*1e3:8.1f} us  {2*B*H*W*9*Ci*Co/ms/1e9:7.1f} TF/s", flush=True)
