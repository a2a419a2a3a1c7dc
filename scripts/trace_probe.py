"""Pipeline timeline of the tensor-core kernel (DF_TC_DBG bit 256): clock64() of cluster 0's producer, issuer and the two stager groups at the
hand-over points of its first 96 k-blocks, for one launch shape.   python scripts/trace_probe.py <l1|l2|tower1|l40>"""
import ctypes, json, os, sys
os.environ["DF_TC_DBG"] = str(int(os.environ.get("DF_TC_DBG", "0")) | 256)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from densefusion_b200 import ops
from densefusion_b200._C import lib
from densefusion_b200.encoder import PackedEncoder, _pack_conv

case = sys.argv[1] if len(sys.argv) > 1 else "l1"
dev = "cuda"
if case == "tower1":
    M, N, K = 64000, 1920, 384
    A = torch.randn(M, K, device=dev); W = ops.SplitWeight(torch.randn(N, K, device=dev) / K ** 0.5); b = torch.randn(128, N, device=dev)
    C = torch.empty(M, N, device=dev)
    run = lambda: ops.gemm(A, W, b, C, M=M, N=N, K=K, lda=K, ldw=K, ldc=N, relu=True, precision="hybrid16s", bias_crop_stride=N, rows_per_crop=500)
else:
    B, H, Wd, ci, co, d = {"l1": (64, 40, 40, 64, 64, 1), "l2": (64, 20, 20, 128, 128, 1), "l40": (64, 20, 20, 512, 512, 1)}[case]
    x = torch.randn(B, H, Wd, ci, device=dev); w = _pack_conv(torch.randn(co, ci, 3, 3, device=dev) / (9 * ci) ** 0.5)
    o = torch.empty(B, H, Wd, co, device=dev)
    run = lambda: PackedEncoder._conv(x, w, o, taps=9, dil=d, act=1, mode=6)
for _ in range(3):
    run()
torch.cuda.synchronize()
EV = 17
buf = (ctypes.c_ulonglong * (EV * 96))()
assert lib.df_tc_trace_read(ctypes.cast(buf, ctypes.c_void_p), EV * 96) == 0
t = np.array(buf, dtype=np.int64).reshape(EV, 96)
names = ["prod_slot_free", "prod_issued", "iss_full", "iss_afull", "iss_done", "stg_full", "stg_split", "stg_slot_free", "stg_handed"]
t0 = t[t > 0].min()
rows = []
for it in range(24, 72):
    rows.append({"it": it, **{n: int(t[e, it] - t0) if t[e, it] else None for e, n in enumerate(names)}})
print(json.dumps({"case": case, "per_kblock_clks": float((t[4, 71] - t[4, 24]) / 47.0)}))
for r in rows[:36]:
    print(" ".join(f"{k}={v}" for k, v in r.items()))

# compact statistics over k-blocks 24..71: mean period of every event and mean latency between consecutive hand-over points
sel = slice(24, 72)
def period(e):
    v = t[e, sel]; v = v[v > 0]
    return float(np.mean(np.diff(v))) if len(v) > 2 else None
print(json.dumps({"period": {n: period(e) for e, n in enumerate(names)},
                  "latency": {"slot_free->issued": float(np.mean(t[1, sel] - t[0, sel])), "issued->stager_sees_bytes": float(np.mean(t[5, sel] - t[1, sel])),
                              "stager: bytes->split": float(np.mean(t[6, sel] - t[5, sel])), "stager: split->slot_free": float(np.mean(t[7, sel] - t[6, sel])),
                              "stager: slot_free->handed": float(np.mean(t[8, sel] - t[7, sel])), "handed->issuer_has_A": float(np.mean(t[3, sel] - t[8, sel])),
                              "issuer: A->committed": float(np.mean(t[4, sel] - t[3, sel])), "issuer: committed->next_loop": float(np.mean(t[2, 25:72] - t[4, 24:71]))}}))

# per accumulation run (tile): the epilogue's and the issuer's view of the accumulator hand-over, and the chunks of epilogue warp 10
ev = ["epi_wait", "epi_has_acc", "epi_drained", "iss_wait_acc", "iss_has_acc"]
for ti in range(0, 10):
    print(f"run={ti} " + " ".join(f"{n}={int(t[9 + e, ti] - t0) if t[9 + e, ti] else None}" for e, n in enumerate(ev)))
for ti in range(1, 7):
    for i in range(3):
        k = ti * 4 + i
        print(f"run={ti} chunk={i} acc_in_regs={int(t[14, k] - t0) if t[14, k] else None} transposed={int(t[15, k] - t0) if t[15, k] else None} stored={int(t[16, k] - t0) if t[16, k] else None}")
