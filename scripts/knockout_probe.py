"""Timing knock-outs of the tensor-core GEMM (DF_TC_DBG bits: 1 no global stores, 2 no transpose, 4 no TMEM load, 8 no MMAs) at the
tower-1 shape over K: which stage of the epilogue / main loop the time follows.  Results of knocked-out runs are wrong by design."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densefusion_b200 import ops
dev = "cuda"
def timeit(fn, reps=20):
    for _ in range(3): fn(0); fn(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i & 1)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {"dbg": os.environ.get("DF_TC_DBG", "0")}
for (M, N, K) in ((64000, 1920, 384), (64000, 1920, 128), (64000, 512, 256)):
    A = [torch.randn(M, K, device=dev) for _ in range(2)]
    W = ops.SplitWeight(torch.randn(N, K, device=dev) / K ** 0.5)
    b = torch.randn(N, device=dev)
    C = [torch.empty(M, N, device=dev) for _ in range(2)]
    def run(i): ops.gemm(A[i], W, b, C[i], M=M, N=N, K=K, lda=K, ldw=K, ldc=N, relu=True, precision="hybrid16s")
    out[f"{M}x{N}x{K}"] = round(timeit(run), 4)
print(json.dumps(out), flush=True)
