#!/usr/bin/env python
"""bench.py -- object poses/sec (estimate + 2 refine iterations) on synthetic YCB-shaped frames.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): PoseNet(num_points=500,
num_obj=21) + 2 PoseRefineNet iterations with eval_ycb semantics, frames of 8 objects (crop mix 3x80^2,
3x120^2, 2x160^2), random-init weights, synthetic inputs.  One step = FRAMES frames (8*FRAMES crops) per GPU;
frames shard across ranks with no data-path collective (weak scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|3xtf32|tf32]

Prints ONE JSON line (rank 0).  `value` = device-resident throughput, `e2e` = through the public API with
pinned host buffers (H2D of every input and D2H of the poses inside the timed region)."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION in this image) off it
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

import numpy as np  # noqa: E402
import torch  # noqa: E402

N_POINTS, N_OBJ, N_MESH, ITERS = 500, 21, 500, 2
CROP_MIX = [(80, 80), (80, 80), (80, 80), (120, 120), (120, 120), (120, 120), (160, 160), (160, 160)]
METRIC = "object poses/sec (estimate + 2 refine iterations)"
UNIT = "poses/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DF_PRECISION", "hybrid16"), choices=["fp32", "3xtf32", "tf32", "hybrid", "hybrid16", "hybrid16p"])
    ap.add_argument("--frames", type=int, default=32, help="frames (of 8 objects) per GPU per step")
    ap.add_argument("--chunk", type=int, default=128, help="crops per head chunk (measured 16: 21.7 ms, 32: 20.7, 64: 20.3, 128: 20.05 per step)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--encoder", default="auto", choices=["auto", "tc", "torch"],
                    help="tc: hand-written tensor-core encoder (default with a tensor-core precision); torch: cuDNN fp32")
    ap.add_argument("--workload", default="pose", choices=["pose", "train"],
                    help="pose: the headline metric; train: config C4, data-parallel training step (16 samples / GPU / step)")
    ap.add_argument("--phase", default="estimator", choices=["estimator", "refiner"])
    ap.add_argument("--num-points", type=int, default=500,
                    help="points per crop: 500 = the configuration the metric is quoted on; 1000 = the reference's YCB setting (config C2 variant)")
    args = ap.parse_args()
    global N_POINTS
    N_POINTS = args.num_points
    return args


def workload_name(frames):
    return (f"YCB PoseNet({N_POINTS},21) + {ITERS} PoseRefineNet iterations (eval_ycb pipeline), {frames} synthetic frames x 8 "
            f"objects per GPU per step, crops 3x80^2+3x120^2+2x160^2, CNN encoder included")


# ------------------------------------------------------------------------------------------------
# synthetic frames (host, pinned)
# ------------------------------------------------------------------------------------------------
def make_host_buckets(frames: int, seed: int, pin: bool):
    """Crops of `frames` frames grouped into (H,W) buckets: list of dicts of host tensors."""
    from densefusion_b200 import synth
    g = torch.Generator().manual_seed(seed)
    buckets = []
    for hw in sorted(set(CROP_MIX)):
        per_frame = CROP_MIX.count(hw)
        b = frames * per_frame
        img = torch.randn(b, 3, hw[0], hw[1], generator=g)
        choose = torch.stack([torch.sort(torch.randperm(hw[0] * hw[1], generator=g)[:N_POINTS])[0] for _ in range(b)])
        cloud = torch.randn(b, N_POINTS, 3, generator=g) * 0.05 + torch.tensor([0.0, 0.0, 0.8])
        obj = torch.randint(0, N_OBJ, (b,), generator=g)
        d = dict(img=img, cloud=cloud, choose=choose.view(b, 1, N_POINTS), obj=obj)
        if pin:
            d = {k: v.pin_memory() for k, v in d.items()}
        buckets.append(d)
    return buckets


def build_modules(device):
    from densefusion_b200 import synth
    from densefusion_b200.lib.network import PoseNet, PoseRefineNet
    est, ref = PoseNet(N_POINTS, N_OBJ), PoseRefineNet(N_POINTS, N_OBJ)
    est_sd = synth.synth_state_dict(synth.shapes_of(est), 0)
    ref_sd = synth.synth_state_dict(synth.shapes_of(ref), 1)
    est.load_state_dict(est_sd)
    ref.load_state_dict(ref_sd)
    est.eval().requires_grad_(False)
    ref.eval().requires_grad_(False)
    if device is not None:
        est.to(device)
        ref.to(device)
    return est, ref, est_sd, ref_sd


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm / baseline: the oracle port of the reference path (torch-CPU fp32 + float64 host algebra)
# ------------------------------------------------------------------------------------------------
def cpu_pose_rate(est_sd, ref_sd, frames: int, warm: int = 1):
    from oracle import df_oracle as O
    buckets = make_host_buckets(frames, seed=999, pin=False)
    crops = [(b["img"][i:i + 1], b["cloud"][i:i + 1], b["choose"][i:i + 1], b["obj"][i].view(1, 1))
             for b in buckets for i in range(b["cloud"].shape[0])]
    for c in crops[:warm]:
        O.estimate_and_refine(est_sd, ref_sd, *c, N_OBJ, ITERS)
    t0 = time.perf_counter()
    for c in crops:
        O.estimate_and_refine(est_sd, ref_sd, *c, N_OBJ, ITERS)
    dt = time.perf_counter() - t0
    return len(crops) / dt, dt, len(crops)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    _, _, est_sd, ref_sd = build_modules(None)
    frames_per_step = 8           # 64 poses ~ 1.4 s of CPU work per step on 16 cores: the default run stays under a minute
    for _ in range(args.warmup):
        cpu_pose_rate(est_sd, ref_sd, frames_per_step, warm=0)
    t0 = time.perf_counter()
    poses = 0
    for _ in range(args.steps):
        _, _, n = cpu_pose_rate(est_sd, ref_sd, frames_per_step, warm=0)
        poses += n
    dt = time.perf_counter() - t0
    val = poses / dt
    sample = f"{frames_per_step} frame x 8 objects per step, {args.steps} steps"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": workload_name(args.frames), "reference_impl": "oracle port (torch-CPU fp32) of the "
                       "reference's Python path; the reference itself is Python and cannot travel to the GPU box"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def time_kernel_ms(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def dominant_kernel_roofline(pipe, precision, peaks):
    """Live timing of the dominant kernel of the step: the first tower layer (conv1_{r,t,c} on the 384 local
    channels, N=1920) over one chunk -- 36.7% of the head's MACs."""
    from densefusion_b200 import engine, ops
    crops, n = pipe.chunk, pipe.n
    rows = crops * n
    ws, _ = pipe._workspaces(crops)
    w = pipe.w_head
    ws.gbias.zero_()
    ws.pf.normal_()

    def run():
        ops.gemm(ws.pf, w.w1_local, ws.gbias, ws.h1, M=rows, N=1920, K=384, lda=384, ldw=384, ldc=1920, relu=True,
                 precision=precision, bias_crop_stride=1920, rows_per_crop=n)
    ms = time_kernel_ms(run)
    flops = 2.0 * rows * 1920 * 384
    if precision == "fp32":
        peak = 148 * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
        peak_note = "fp32 FFMA pipe, 148 SM x 128 lanes x 2 x max SM clock (no measured fp32 figure in MEASURED_PEAKS.json)"
        bound = "tensor"
        kname = "sgemm_kernel<128,128> (fp32 FFMA)"
    else:
        peak = peaks.get("bf16_tflops", 1590.0) / 2.0
        peak_note = ("TF32 tcgen05 peak taken as half of the measured bf16 burst figure of MEASURED_PEAKS.json"
                     if "bf16_tflops" in peaks else "TF32 = half of the fallback 1.59 PFLOP/s bf16")
        bound = "tensor"
        kinds = {"hybrid": "kind::tf32 + kind::f16 bf16 corrections", "hybrid16": "kind::f16: fp16 main term + bf16 corrections"}
        kname = "gemm_tc_q_kernel<2,2> (tcgen05.mma.cta_group::2 %s, TMA operands, %s)" % (kinds.get(precision, "kind::tf32"), precision)
    ach = flops / (ms * 1e-3) / 1e12
    shape = f"M={rows} N=1920 K=384"
    # what the library reaches on this GPU in single-pass TF32 (cuBLAS, 8192^3, best of 10): a measured TF32 ceiling next to
    # the derived one (MEASURED_PEAKS.json has no TF32 figure)
    tf32_lib = None
    if precision != "fp32":
        try:
            a = torch.randn(8192, 8192, device=ws.pf.device)
            b = torch.randn(8192, 8192, device=ws.pf.device)
            torch.backends.cuda.matmul.allow_tf32 = True
            best = min(time_kernel_ms(lambda: torch.matmul(a, b), iters=1, warm=1) for _ in range(10))
            tf32_lib = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
        except Exception:
            tf32_lib = None
        finally:
            torch.backends.cuda.matmul.allow_tf32 = False
    traffic = None
    try:
        if precision != "fp32":
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))[shape]["bytes"]
    except Exception:
        traffic = None
    return {"bound": bound, "kernel": kname, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": traffic, "ms_per_launch": ms, "algorithmic_flops_per_launch": flops,
            "frac_of_mode_ceiling": ({"3xtf32": 3.0, "hybrid": 2.0, "hybrid16": 1.5}.get(precision, 1.0) * ach / peak) if precision != "fp32" else None,
            "l2": "operands + output of one launch (596 MB at the default chunk) exceed the 126 MB L2",
            "cublas_tf32_8192_tflops": tf32_lib, "frac_vs_cublas_tf32": (ach / tf32_lib) if tf32_lib else None,
            "executed_over_algorithmic": {"3xtf32": 3.0, "hybrid": 2.0, "hybrid16": 1.5}.get(precision, 1.0), "peak_source": peak_note,
            "shape": shape}


def run_ours(args):
    from densefusion_b200 import _C, ops, synth
    from densefusion_b200.pipeline import GraphedBuckets, PoseEstimator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference "
                         "for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False          # the encoder runs in true fp32 (parity mode)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = os.environ.get("DF_CUDNN_BENCHMARK", "1") == "1"

    est, ref, est_sd, ref_sd = build_modules(dev)
    pipe = PoseEstimator(est, ref, iterations=ITERS, precision=args.precision, chunk_crops=args.chunk, encoder=args.encoder)
    # two distinct input sets per rank, alternated between steps
    host_sets = [make_host_buckets(args.frames, seed=1000 + 17 * rank + s, pin=True) for s in range(2)]
    dev_sets = [[{k: v.to(dev) for k, v in b.items()} for b in hs] for hs in host_sets]
    shapes = [(b["cloud"].shape[0], b["img"].shape[2], b["img"].shape[3]) for b in host_sets[0]]
    crops_per_step = sum(s[0] for s in shapes)

    launch_mode = "stream"
    graphed = None
    if not args.no_graph:
        try:
            graphed = GraphedBuckets(pipe, shapes)
            launch_mode = "cuda_graph"
        except Exception as e:       # capture problems only change HOW kernels are launched, not what runs
            graphed = None
            launch_mode = f"stream (graph capture failed: {type(e).__name__})"
            torch.cuda.synchronize()

    pose_host = torch.empty(crops_per_step, 7, dtype=torch.float64).pin_memory()
    streaming = None
    if graphed is not None:
        try:
            from densefusion_b200.pipeline import StreamingEstimator
            streaming = StreamingEstimator(pipe, shapes)
        except Exception:
            streaming = None
            torch.cuda.synchronize()
    pending = []

    def preload_device_sets():
        """`value`: inputs already resident in HBM -- the two input sets sit in the static buffers of the two graphs."""
        if streaming is not None:
            for slot, ds in zip(streaming.slots, dev_sets):
                for s, d in zip(slot.static, ds):
                    for k in ("img", "cloud", "choose", "obj"):
                        s[k].copy_(d[k].view(s[k].shape))
            torch.cuda.synchronize()

    def step_device(i):
        if streaming is not None:
            return streaming.slots[i & 1].run()
        if graphed is not None:
            for s, d in zip(graphed.static, dev_sets[i & 1]):
                for k in ("img", "cloud", "choose", "obj"):
                    s[k].copy_(d[k])
            return graphed.run()
        return pipe.estimate_buckets(dev_sets[i & 1])

    def step_e2e(i):
        """Public serving API with pinned host inputs: every step copies its inputs H2D and its poses D2H.  With the
        streaming estimator the copy of step i+1 overlaps the compute of step i and the host reads the poses of step
        i-1 (one step of latency, every result is read); `e2e_drain` waits for the last one inside the timed region."""
        hs = host_sets[i & 1]
        if streaming is not None:
            pending.append(streaming.submit(hs))
            if len(pending) > 1:
                return streaming.result(pending.pop(0))
            return None
        out = pipe.estimate_buckets([{k: v.to(dev, non_blocking=True) for k, v in b.items()} for b in hs])
        pose_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the user reads the poses every step
        return pose_host

    def e2e_drain():
        while pending:
            streaming.result(pending.pop(0))

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup, drain=None):
        for i in range(warmup):
            step_fn(i)
        if drain:
            drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            step_fn(i)
        if drain:
            drain()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall

    preload_device_sets()
    # sanity: the timed pipeline produces finite unit quaternions
    chk = step_device(0).cpu()
    assert torch.isfinite(chk).all() and torch.allclose(chk[:, :4].norm(dim=1), torch.ones(crops_per_step, dtype=torch.float64), atol=1e-6)

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _C.lib.launches
    ms, _ = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop()
    launches_eager = _C.lib.launches - launches0
    value = world * crops_per_step * args.steps / (ms * 1e-3)

    ms_e2e, wall_e2e = timed(step_e2e, args.steps, args.warmup, e2e_drain)
    e2e_val = world * crops_per_step * args.steps / (max(ms_e2e * 1e-3, wall_e2e))
    h2d = sum(v.numel() * v.element_size() for b in host_sets[0] for v in b.values())
    d2h = pose_host.numel() * pose_host.element_size()

    # kernels of OURS per step (counted once in eager mode so the number is exact even when graphs replay)
    l0 = _C.lib.launches
    pipe.estimate_buckets(dev_sets[0])
    torch.cuda.synchronize()
    launches_per_step = _C.lib.launches - l0

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        roof = dominant_kernel_roofline(pipe, args.precision, peaks)
        extras = {}
        if not args.no_extras:
            extras = measure_extras(pipe, dev, args, peaks)
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            torch.set_num_threads(os.cpu_count() or 1)
            rate, dt, n = cpu_pose_rate(est_sd, ref_sd, 60)
            cpu = {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"60 frames x 8 objects = {n} poses of the same crop mix, {dt:.1f} s, oracle port of the reference "
                             "path (CNN + head + select + 2 refine iterations, torch-CPU fp32, all host threads)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": {"fp32": "fp32", "hybrid": "fp32 (fp32-parity tensor-core GEMMs: TF32 main term + bf16 correction terms, fp32 accumulate)",
                          "hybrid16": "fp32 (fp32-parity tensor-core GEMMs: fp16 main term + bf16 correction terms, fp32 accumulate)",
                          "hybrid16p": "fp32 (fp32-parity tensor-core GEMMs: fp16 main term + bf16 correction terms, fp32 accumulate)",
                          "3xtf32": "fp32 (fp32-parity tensor-core GEMMs: 3xTF32, fp32 accumulate)",
                          "tf32": "tf32 (single-pass tensor-core GEMMs, fp32 accumulate; looser bound)"}[args.precision],
                "data": "synthetic",
                "config": {"workload": workload_name(args.frames), "num_points": N_POINTS, "num_obj": N_OBJ,
                           "refine_iterations": ITERS, "crops_per_gpu_per_step": crops_per_step,
                           "precision": args.precision,
                           "encoder": ("densefusion_b200.encoder: tcgen05 implicit-GEMM convolutions, NHWC, " + args.precision)
                           if pipe.encoder == "tc" else "torch/cuDNN strict fp32 (TF32 off), NCHW",
                           "launch": launch_mode, "chunk_crops": args.chunk,
                           "l2": "two alternating input sets; per-step working set (encoder activations > 1 GB) exceeds the 126 MB L2",
                           "parallelism": f"frames sharded over {world} GPU(s), no data-path collective"},
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps,
                        "api": "pipeline.StreamingEstimator (double-buffered H2D on a copy stream)" if streaming is not None
                               else "pipeline.PoseEstimator.estimate_buckets"},
                "gpu_launches": launches_per_step * args.steps,
                "gpu_launches_per_step": launches_per_step,
                "roofline": roof}
        if cpu:
            line["cpu_baseline"] = cpu
        if extras:
            line["extras"] = extras
    barrier()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# config C4: data-parallel training step (tools/train.py:143-169), 16 samples per GPU per optimiser step
# ------------------------------------------------------------------------------------------------
TRAIN_MIX = [(80, 80)] * 6 + [(120, 120)] * 6 + [(160, 160)] * 4


def make_train_buckets(seed: int, pin: bool):
    from densefusion_b200 import synth
    g = torch.Generator().manual_seed(seed)
    buckets = []
    for hw in sorted(set(TRAIN_MIX)):
        b = TRAIN_MIX.count(hw)
        img = torch.randn(b, 3, hw[0], hw[1], generator=g)
        choose = torch.stack([torch.sort(torch.randperm(hw[0] * hw[1], generator=g)[:N_POINTS])[0] for _ in range(b)])
        points = torch.randn(b, N_POINTS, 3, generator=g) * 0.05 + torch.tensor([0.0, 0.0, 0.8])
        model = torch.randn(b, N_MESH, 3, generator=g) * 0.05
        rot = torch.stack([synth.quat_to_rot(synth.random_unit_quaternion(g)) for _ in range(b)])
        target = torch.bmm(model, rot.transpose(1, 2)) + torch.tensor([0.0, 0.0, 0.8])
        idx = torch.randint(0, N_OBJ, (b, 1), generator=g)
        d = dict(img=img, points=points, choose=choose.view(b, 1, N_POINTS), idx=idx, target=target.contiguous(),
                 model_points=model)
        if pin:
            d = {k: v.pin_memory() for k, v in d.items()}
        buckets.append(d)
    return buckets


def measure_train(dev, phase, steps, warmup, world, rank, barrier, no_graph=False):
    """ms per optimiser step (max over ranks done by the caller) with device-resident inputs and end to end."""
    from densefusion_b200 import synth
    from densefusion_b200.trainer import DataParallelTrainer, GraphedTrainStep
    est, ref, _, _ = build_modules(dev)
    tr = DataParallelTrainer(est, ref, N_MESH, synth.YCB_SYM, lr=1e-4, w=0.015, iteration=ITERS, phase=phase)
    host_sets = [make_train_buckets(7000 + 31 * rank + s, pin=True) for s in range(2)]
    dev_sets = [[{k: v.to(dev) for k, v in b.items()} for b in hs] for hs in host_sets]
    samples = sum(b["points"].shape[0] for b in host_sets[0])
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
    graphed, launch = None, "stream"
    if not no_graph:
        try:
            graphed = GraphedTrainStep(tr, dev_sets[0])
            launch = "cuda_graph"
        except Exception as e:
            graphed, launch = None, f"stream (graph capture failed: {type(e).__name__}: {e})"[:200]
            torch.cuda.synchronize()

    def step_device(i):
        return graphed.step(dev_sets[i & 1]) if graphed is not None else tr.step(dev_sets[i & 1])

    def step_e2e(i):
        if graphed is not None:
            out = graphed.step(host_sets[i & 1])
        else:
            out = tr.step([{k: v.to(dev, non_blocking=True) for k, v in b.items()} for b in host_sets[i & 1]])
        loss_host.copy_(out["loss_sum"].reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def timed(fn):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / steps

    ms_dev = timed(step_device)
    ms_e2e = timed(step_e2e)
    h2d = sum(v.numel() * v.element_size() for b in host_sets[0] for v in b.values())
    arena = tr.arena_est if phase == "estimator" else tr.arena_ref
    return {"ms_per_step": ms_dev, "ms_per_step_e2e": ms_e2e, "samples_per_gpu": samples, "h2d_bytes_per_step": h2d,
            "allreduce_bytes": arena.total * 4, "parameters": arena.numel, "launch": launch}


def run_train(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --workload train: no CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = os.environ.get("DF_CUDNN_BENCHMARK", "1") == "1"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    r = measure_train(dev, args.phase, args.steps, args.warmup, world, rank, barrier, args.no_graph)
    clocks = sampler.stop()
    t = torch.tensor([r["ms_per_step"], r["ms_per_step_e2e"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    total = world * r["samples_per_gpu"]
    if rank == 0:
        print(json.dumps({
            "metric": "training samples/sec (data-parallel step: forward, fused loss, backward, all-reduce SUM, Adam)",
            "value": total / (ms * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
            "data": "synthetic",
            "config": {"workload": f"config C4: YCB PoseNet({N_POINTS},21) {args.phase} phase, 16 samples per GPU per optimiser step "
                                   "(6x80^2 + 6x120^2 + 4x160^2), gradient SUM semantics of tools/train.py:159-169",
                       "global_batch": total, "phase": args.phase, "allreduce_bytes": r["allreduce_bytes"],
                       "parameters": r["parameters"], "launch": r["launch"],
                       "parallelism": f"dp{world}, one NCCL all-reduce per step"},
            "clocks": clocks,
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": r["h2d_bytes_per_step"],
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e}}))
    barrier()
    if world > 1:
        dist.destroy_process_group()


def measure_extras(pipe, dev, args, peaks):
    """Per-stage numbers that explain the headline: head-only poses/s, and config C1 (ADD-S loss, 256 crops)."""
    from densefusion_b200 import ops, synth
    out = {}
    B, n = 256, N_POINTS
    g = torch.Generator().manual_seed(5)
    cloud = (torch.randn(B, n, 3, generator=g) * 0.05 + torch.tensor([0.0, 0.0, 0.8])).to(dev)
    emb_pm = torch.log_softmax(torch.randn(B * n, 32, generator=g), dim=1).to(dev)
    obj = torch.randint(0, N_OBJ, (B,), generator=g).to(dev)
    ms = time_kernel_ms(lambda: pipe.head_and_refine(cloud, emb_pm, obj), iters=5, warm=2)
    out["head_plus_refine_only"] = {"value": B / (ms * 1e-3), "unit": UNIT, "crops": B,
                                    "note": "K1-K5 only: embeddings precomputed, encoder excluded"}
    ms0 = time_kernel_ms(lambda: pipe.head_and_refine(cloud, emb_pm, obj, iterations=0), iters=5, warm=2)
    out["head_only_ms_per_256_crops"] = ms0
    alg_flops = B * (2.002112e6 * n + 1.96608e6) * 2
    out["head_algorithmic_tflops"] = alg_flops / (ms0 * 1e-3) / 1e12
    # config C1: ADD-S loss (kNN R=500, Q=250000 per crop) over 256 crops
    pr = torch.randn(B, n, 4, generator=g).to(dev)
    pt = (torch.randn(B, n, 3, generator=g) * 0.02).to(dev)
    pc = (torch.rand(B, n, 1, generator=g) * 0.9 + 0.05).to(dev)
    model = (torch.randn(B, N_MESH, 3, generator=g) * 0.05).to(dev)
    target = model + torch.tensor([0.0, 0.0, 0.8], device=dev)
    sym_obj = torch.full((B,), 12, dtype=torch.int64, device=dev)
    mask = ops.sym_mask(synth.YCB_SYM)
    ms_s = time_kernel_ms(lambda: ops.loss_forward(pr, pt, pc, target, model, cloud, cloud, sym_obj, mask, True, 0.015), iters=5, warm=2)
    ms_a = time_kernel_ms(lambda: ops.loss_forward(pr, pt, pc, target, model, cloud, cloud, sym_obj, mask, False, 0.015), iters=5, warm=2)
    pairs = B * n * N_MESH * N_MESH
    out["c1_adds_loss_256_crops"] = {"ms": ms_s, "crops_per_s": B / (ms_s * 1e-3), "pair_evals_per_s": pairs / (ms_s * 1e-3),
                                     "algorithmic_bytes": B * 46e3, "hbm_gbs": B * 46e3 / (ms_s * 1e-3) / 1e9,
                                     "fp32_lane_ops_per_s": pairs * 9 / (ms_s * 1e-3)}
    out["c1_add_loss_256_crops"] = {"ms": ms_a, "crops_per_s": B / (ms_a * 1e-3), "hbm_gbs": B * 46e3 / (ms_a * 1e-3) / 1e9}
    # the same whole pipeline with the encoder allowed to use cuDNN's TF32 channels-last kernels (NOT the fp32-parity
    # configuration: embeddings then differ at the 1e-3 level) -- shows how much of the step is the library encoder
    import copy
    from densefusion_b200.pipeline import PoseEstimator
    try:
        est2 = copy.deepcopy(pipe.estimator)
        pipe2 = PoseEstimator(est2, pipe.refiner, iterations=ITERS, precision=pipe.precision, chunk_crops=pipe.chunk,
                              channels_last=True, encoder="torch")
        buckets = [{k: v.to(dev) for k, v in b.items()} for b in make_host_buckets(args.frames, seed=4242, pin=False)]
        torch.backends.cudnn.allow_tf32 = True
        ms_t = time_kernel_ms(lambda: pipe2.estimate_buckets(buckets), iters=3, warm=2)
        crops = sum(b["cloud"].shape[0] for b in buckets)
        out["whole_pipeline_with_cudnn_tf32_nhwc_encoder"] = {
            "value": crops / (ms_t * 1e-3), "unit": UNIT, "ms_per_step": ms_t,
            "note": "library encoder (torch/cuDNN channels-last) with TF32 allowed: NOT a parity mode (embeddings off by ~1e-3); "
                    "for comparison with the hand-written fp32-parity encoder"}
        torch.backends.cudnn.allow_tf32 = False
        pipe3 = PoseEstimator(copy.deepcopy(pipe.estimator), pipe.refiner, iterations=ITERS, precision=pipe.precision,
                              chunk_crops=pipe.chunk, encoder="torch")
        ms_c = time_kernel_ms(lambda: pipe3.estimate_buckets(buckets), iters=3, warm=2)
        out["whole_pipeline_with_cudnn_fp32_encoder"] = {"value": crops / (ms_c * 1e-3), "unit": UNIT, "ms_per_step": ms_c,
                                                         "note": "library encoder in strict fp32 (the round-1 default before the tensor-core encoder)"}
    finally:
        torch.backends.cudnn.allow_tf32 = False
    return out


if __name__ == "__main__":
    a = parse()
    # stdout carries exactly ONE JSON line: while the bench runs, file descriptor 1 points at stderr, so anything a library
    # writes to the process's stdout (NCCL prints its version banner there at every debug level >= VERSION) cannot precede it;
    # Python's own print() goes to the real stdout through a separate handle.
    sys.stdout.flush()
    _real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_real, "w", buffering=1)
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "train":
        run_train(a)
    else:
        run_ours(a)
    sys.stdout.flush()
