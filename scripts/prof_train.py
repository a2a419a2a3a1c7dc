"""Per-kernel time of ONE eager training step (torch.profiler / CUPTI, no replay): which kernels the step is made of."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from densefusion_b200 import synth
from densefusion_b200.trainer import DataParallelTrainer

phase = sys.argv[1] if len(sys.argv) > 1 else "estimator"
dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
est, ref, _, _ = bench.build_modules(dev)
tr = DataParallelTrainer(est, ref, bench.N_MESH, synth.YCB_SYM, phase=phase)
buckets = [{k: v.to(dev) for k, v in b.items()} for b in bench.make_train_buckets(1, pin=False)]
for _ in range(3):
    tr.step(buckets)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(buckets)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
total = sum(e.device_time_total for e in rows)
print(f"# phase {phase}: {len(rows)} distinct kernels, total device time {total/1e3:.2f} ms")
for e in rows[:70]:
    print(f"{e.device_time_total/1e3:9.3f} ms {100*e.device_time_total/total:5.1f}%  x{e.count:<4d} {e.key[:110]}")
