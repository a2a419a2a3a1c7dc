"""Per-kernel device time of ONE eager inference step of the bench workload (torch.profiler / CUPTI)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from densefusion_b200.pipeline import PoseEstimator

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 32
precision = sys.argv[2] if len(sys.argv) > 2 else "hybrid16"
dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
est, ref, _, _ = bench.build_modules(dev)
pipe = PoseEstimator(est, ref, iterations=2, precision=precision, chunk_crops=256)
buckets = [{k: v.to(dev) for k, v in b.items()} for b in bench.make_host_buckets(frames, 3, pin=False)]
for _ in range(3):
    pipe.estimate_buckets(buckets)
torch.cuda.synchronize()
if os.environ.get("DF_NCU") == "1":
    # under `ncu --profile-from-start off --metrics gpu__time_duration.sum`: exactly one step between the profiler marks
    torch.cuda.cudart().cudaProfilerStart()
    pipe.estimate_buckets(buckets)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    sys.exit(0)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pipe.estimate_buckets(buckets)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
total = sum(e.device_time_total for e in rows)
print(f"# {frames} frames x 8 crops, encoder={pipe.encoder}: {len(rows)} distinct kernels, total device time {total/1e3:.2f} ms")
for e in rows[:40]:
    print(f"{e.device_time_total/1e3:9.3f} ms {100*e.device_time_total/total:5.1f}%  x{e.count:<4d} {e.key[:100]}")
