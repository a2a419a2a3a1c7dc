"""Config C3: ADD-S kNN sweep, R = M model points x P per-pixel hypotheses (Q = P*M queries), k = 1, D = 3.
Throughput of df_knn (pair evaluations / s) and a bit-exactness spot check against the reference's own CUDA kernel
(oracle/_ref/libknn_reference.so, query chunked) on the first 200 000 queries of every case."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from densefusion_b200 import ops
from util import reference_knn_gpu

rows = []
g = torch.Generator().manual_seed(0)
for R in (500, 1000, 2600, 5000, 10000, 20000):
    ref = (torch.randn(1, 3, R, generator=g) * 0.05).cuda()
    for P in (500, 1000, 2048, 4096):
        Q = P * R
        qry = (torch.randn(1, 3, min(Q, 1 << 22), generator=g) * 0.05).cuda()
        qry = qry.repeat(1, 1, (Q + qry.shape[2] - 1) // qry.shape[2])[:, :, :Q].contiguous()
        out = torch.empty(1, 1, Q, dtype=torch.int64, device="cuda")
        ops.knn(ref, qry, 1, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.knn(ref, qry, 1, out=out); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        nchk = min(Q, 200000)
        want = reference_knn_gpu(ref[0], qry[0, :, :nchk].contiguous(), 1)
        exact = None if want is None else bool(torch.equal(out[0, :, :nchk], want))
        rows.append({"R": R, "P": P, "Q": Q, "ms": round(ms, 3), "pairs_per_s": R * Q / (ms * 1e-3),
                     "bytes_GBps": (12 * R + 20 * Q) / (ms * 1e-3) / 1e9, "bit_exact_vs_reference_kernel": exact})
        print(json.dumps(rows[-1]), flush=True)
        del qry, out
print(json.dumps({"summary": "df_knn C3 sweep", "cases": len(rows), "all_exact": all(r["bit_exact_vs_reference_kernel"] in (True, None) for r in rows)}))
