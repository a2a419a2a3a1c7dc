"""Data-parallel training step of the pose path (BASELINE.json config C4; reference: tools/train.py:131-169).

The reference trains with batch size 1 and *accumulates* un-normalised gradients over `batch_size` samples before one
Adam step (tools/train.py:159-169).  Here a rank evaluates its crops batched per (H,W) bucket, every crop's loss gets
gradient 1 (sum semantics), and the ranks exchange gradients exactly once per optimiser step:

    flat fp32 arena  [ parameters | gradients | exp_avg | exp_avg_sq ]      (views alias module.parameters())
        backward kernels write / accumulate straight into the gradient arena
        ONE all-reduce(SUM) of the gradient arena   (NCCL over NVLink / NVSwitch; estimator 85.8 MB, refiner 7.7 MB)
        ONE df_adam_step launch over the parameter arena (torch.optim.Adam arithmetic, identical on every rank)

Two phases as in the reference: "estimator" optimises all of PoseNet (tools/train.py:97), "refiner" runs PoseNet
without a graph and optimises PoseRefineNet through `iteration` chained Loss_refine evaluations, each with its own
backward (tools/train.py:156-159).  Inference needs none of this: frames shard with no collective (`shard_range`)."""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch

from ._C import DFError, check, lib, ptr, stream
from .lib import conv_tc

ALIGN = 64          # floats; keeps every parameter view 256-byte aligned (float4 / TMA-able)


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous block partition of n_items over `world` ranks (first n_items % world ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


class FlatArena:
    """Parameters, gradients and Adam moments of a list of nn.Parameters as four flat fp32 buffers.
    After construction p.data and p.grad are views into the arena, so autograd accumulates in place."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        if not self.params:
            raise ValueError("FlatArena: no parameters")
        dev = self.params[0].device
        self.offsets, off = [], 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatArena: fp32 parameters on one device expected")
            self.offsets.append(off)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.total = off
        self.numel = sum(p.numel() for p in self.params)
        z = lambda: torch.zeros(self.total, device=dev, dtype=torch.float32)
        self.param, self.grad, self.exp_avg, self.exp_avg_sq = z(), z(), z(), z()
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)      # optimiser step number, device resident
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                n = p.numel()
                self.param[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.param[o:o + n].view(p.shape)
                p.grad = self.grad[o:o + n].view(p.shape)

    # ---- gradient exchange overlapped with the backward pass ---------------------------------------------------------
    def enable_overlap(self, bucket_bytes: int = 16 << 20) -> None:
        """Cut the gradient arena into contiguous communication buckets (>= bucket_bytes each, from the END of the arena
        backwards: autograd produces the head's gradients first and walks the encoder from its last layer to its first, i.e.
        roughly in reverse parameter order) and hook every parameter: once all parameters of a bucket have received the last of
        the step's `expected` gradient accumulations, that slice is all-reduced asynchronously while the backward kernels of the
        earlier layers are still running.  `finish_all_reduce` launches whatever is left (parameters that never get a gradient,
        e.g. the unused classifier) and makes the current stream wait for every exchange."""
        if getattr(self, "_buckets", None) is not None:
            return
        bounds, end, acc = [], self.total, 0
        for p, o in reversed(list(zip(self.params, self.offsets))):
            acc = end - o
            if acc * 4 >= bucket_bytes:
                bounds.append((o, end))
                end = o
        if end > 0:
            bounds.append((0, end))
        self._buckets = bounds                                    # [(start, stop)] in backward order
        self._bucket_of, self._bucket_size = {}, [0] * len(bounds)
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            b = next(k for k, (lo, hi) in enumerate(bounds) if lo <= o < hi)
            self._bucket_of[i] = b
            self._bucket_size[b] += 1
        self._pending, self._launched, self._works, self._group, self._armed = [], [], [], None, False
        self._fired, self._fired_known = set(), False              # parameters that received a gradient in the first armed step
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _make_hook(self, i):
        def hook(_param):
            if not self._armed:
                return
            if not self._fired_known:
                self._fired.add(i)
            elif i not in self._fired:
                return
            b = self._bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b):
        import torch.distributed as dist
        lo, hi = self._buckets[b]
        self._launched[b] = True
        self._works.append(dist.all_reduce(self.grad[lo:hi], op=dist.ReduceOp.SUM, group=self._group, async_op=True))

    def begin_step(self, expected: int, group=None) -> bool:
        """Arm the overlap for one optimiser step in which every parameter receives `expected` gradient accumulations."""
        import torch.distributed as dist
        if getattr(self, "_buckets", None) is None or not (dist.is_available() and dist.is_initialized()
                                                          and dist.get_world_size(group) > 1):
            return False
        if self._fired_known:                                      # parameters without a gradient (the unused classifier) do not count
            sizes = [0] * len(self._buckets)
            for i in self._fired:
                sizes[self._bucket_of[i]] += 1
        else:
            sizes = self._bucket_size
        self._pending = [n * expected if n else -1 for n in sizes]
        self._launched = [False] * len(self._buckets)
        self._works, self._group, self._armed = [], group, True
        return True

    def finish_all_reduce(self) -> None:
        self._armed = False
        if not self._fired_known and self._fired:
            self._fired_known = True
        for b in range(len(self._buckets)):
            if not self._launched[b]:
                self._launch(b)
        for w in self._works:
            w.wait()                                              # current stream waits for NCCL's stream
        self._works = []

    def zero_grad(self) -> None:
        self.grad.zero_()
        for p, o in zip(self.params, self.offsets):          # re-attach if someone set .grad = None
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def all_reduce(self, group=None) -> None:
        """SUM over ranks (the reference's accumulation is un-normalised, tools/train.py:159-169)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)

    def mark_updated(self) -> None:
        """The arena was written through a raw pointer (optimiser kernel, graph replay, load_state): neither data_ptr() nor
        ._version of the parameters changes by itself, so bump the versions -- engine.param_version keys the packed-weight
        caches of PoseNet / PoseRefineNet / the encoder on them, and a stale cache would evaluate old weights."""
        torch.autograd.graph.increment_version(self.params)

    def adam_step(self, lr: float, betas=(0.9, 0.999), eps: float = 1e-8) -> None:
        if not self.param.is_cuda:
            raise DFError("FlatArena.adam_step: the optimiser step is a CUDA kernel (no CPU fallback)")
        check(lib.df_adam_step_dev(ptr(self.param), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.total,
                                   float(lr), float(betas[0]), float(betas[1]), float(eps), ptr(self.step_dev), stream()),
              "df_adam_step_dev")
        self.mark_updated()

    def reset_optimizer(self) -> None:
        """A fresh Adam: moments and step count back to zero (the reference builds a new optimiser on every schedule switch)."""
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.step_dev.zero_()

    def state(self) -> dict:
        return {k: getattr(self, k).clone() for k in ("param", "exp_avg", "exp_avg_sq", "step_dev")}

    def load_state(self, st: dict) -> None:
        for k, v in st.items():
            getattr(self, k).copy_(v)
        self.mark_updated()


class DataParallelTrainer:
    """One optimiser step per call; every rank passes its own shard of the global batch.

    buckets: list of dicts of batched CUDA tensors sharing one crop size:
        img (b,3,H,W), points (b,N,3), choose (b,1,N), idx (b,1)|(b,), target (b,M,3), model_points (b,M,3)."""

    def __init__(self, estimator, refiner, num_points_mesh: int, sym_list: Sequence[int], lr: float = 1e-4,
                 w: float = 0.015, iteration: int = 2, phase: str = "estimator", group=None,
                 frozen_precision: str = "hybrid16s", overlap: Optional[bool] = None):
        from .lib.loss import Loss
        from .lib.loss_refiner import Loss_refine
        self.estimator, self.refiner = estimator, refiner
        self.criterion = Loss(num_points_mesh, sym_list)
        self.criterion_refine = Loss_refine(num_points_mesh, sym_list)
        self.lr, self.w, self.iteration, self.group = float(lr), float(w), int(iteration), group
        # arithmetic of the FROZEN estimator in the refiner phase (no gradient flows through it, tools/train.py:93):
        # an fp32-parity tensor-core mode runs its encoder + head 4-5x faster than the exact-fp32 / cuDNN path
        self.frozen_precision = frozen_precision
        # overlap the gradient exchange with the backward pass (bucketed asynchronous all-reduces, see FlatArena.enable_overlap);
        # DF_DP_OVERLAP=0 / overlap=False: ONE all-reduce of the whole arena after the last backward kernel (round-1 behaviour)
        import os
        self.overlap = (os.environ.get("DF_DP_OVERLAP", "1") != "0") if overlap is None else bool(overlap)
        self.exchange = os.environ.get("DF_DP_EXCHANGE", "1") != "0"     # 0: skip the exchange (measurement of its cost only)
        self.arena_est: Optional[FlatArena] = None
        self.arena_ref: Optional[FlatArena] = None
        self.set_phase(phase)

    def set_phase(self, phase: str) -> None:
        if phase not in ("estimator", "refiner"):
            raise ValueError("phase must be 'estimator' or 'refiner'")
        self.phase = phase
        if phase == "estimator":
            self.estimator.requires_grad_(True)
            self.estimator.precision = "fp32"
            if self.arena_est is None:
                self.arena_est = FlatArena(self.estimator.parameters())
        else:
            self.estimator.requires_grad_(False)        # tools/train.py:93,228: only the refiner is optimised
            self.estimator.precision = self.frozen_precision
            self.refiner.requires_grad_(True)
            if self.arena_ref is None:
                self.arena_ref = FlatArena(self.refiner.parameters())

    def reset_optimizer(self) -> None:
        arena = self.arena_est if self.phase == "estimator" else self.arena_ref
        arena.reset_optimizer()

    # ---- local forward / backward (gradients land in the arena) ----
    def _local_estimator(self, buckets):
        loss_sum = dis_sum = 0.0
        for b in buckets:
            r, t, c, _ = self.estimator.forward_batched(b["img"], b["points"], b["choose"], b["idx"])
            loss, dis, _, _ = self.criterion(r, t, c, b["target"], b["model_points"], b["idx"], b["points"], self.w, False)
            loss.sum().backward()
            loss_sum = loss_sum + loss.detach().sum()
            dis_sum = dis_sum + dis.detach().sum()
        return loss_sum, dis_sum

    def _local_refiner(self, buckets):
        loss_sum = dis_sum = 0.0
        for b in buckets:
            with torch.no_grad():
                r, t, c, emb = self.estimator.forward_batched(b["img"], b["points"], b["choose"], b["idx"])
                loss, dis, new_points, new_target = self.criterion(r, t, c, b["target"], b["model_points"], b["idx"],
                                                                   b["points"], self.w, True)
            for _ in range(self.iteration):
                pr, pt = self.refiner.forward_batched(new_points, emb, b["idx"])
                dis, new_points, new_target = self.criterion_refine(pr, pt, new_target, b["model_points"], b["idx"],
                                                                    new_points)
                dis.sum().backward()
            loss_sum = loss_sum + loss.detach().sum()
            dis_sum = dis_sum + dis.detach().sum()
        return loss_sum, dis_sum

    def step(self, buckets) -> dict:
        arena = self.arena_est if self.phase == "estimator" else self.arena_ref
        arena.zero_grad()
        fn = self._local_estimator if self.phase == "estimator" else self._local_refiner
        overlapped = False
        if self.overlap and self.exchange:
            arena.enable_overlap()
            backwards = len(buckets) * (1 if self.phase == "estimator" else self.iteration)
            overlapped = arena.begin_step(backwards, self.group)
        with conv_tc.weight_cache():             # every convolution weight is packed once per step, not once per bucket
            loss_sum, dis_sum = fn(buckets)
        if overlapped:
            arena.finish_all_reduce()
        elif self.exchange:
            arena.all_reduce(self.group)
        arena.adam_step(self.lr)
        return {"loss_sum": loss_sum, "dis_sum": dis_sum}


class GraphedTrainStep:
    """CUDA-graph capture of DataParallelTrainer.step for fixed bucket shapes: forward, fused loss, the explicit
    backward kernels, cuDNN's encoder backward, the NCCL all-reduce and the Adam kernel replay as ONE graph launch
    (a training step is ~1500 small launches; eager it is bound by the Python / launch path, not by the GPU).
    Optimiser state is snapshotted around the warm-up so capturing does not advance training."""

    def __init__(self, trainer: DataParallelTrainer, example_buckets, warmup: int = 3):
        self.trainer = trainer
        self.static = [{k: v.clone() for k, v in b.items()} for b in example_buckets]
        arena = trainer.arena_est if trainer.phase == "estimator" else trainer.arena_ref
        saved = arena.state()
        dev = arena.param.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                trainer.step(self.static)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = trainer.step(self.static)
        arena.load_state(saved)
        arena.zero_grad()

    def load(self, buckets) -> None:
        for s, b in zip(self.static, buckets):
            for k in s:
                s[k].copy_(b[k], non_blocking=True)

    def step(self, buckets=None) -> dict:
        if buckets is not None:
            self.load(buckets)
        self.graph.replay()
        arena = self.trainer.arena_est if self.trainer.phase == "estimator" else self.trainer.arena_ref
        arena.mark_updated()                 # the replayed Adam kernel changed the weights behind autograd's back
        return self.out
