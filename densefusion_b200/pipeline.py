"""Batched, on-device estimate -> select -> (re-express -> refine -> compose) x iters pipeline
(subsystem (4) of BASELINE.json; reference: the per-object loop of tools/eval_ycb.py:147-233).

The reference handles one object at a time and leaves the GPU 2 + 2*iters times per object.  Here a
batch of crops (several frames' objects, grouped into (H,W) buckets for the encoder) goes through the
whole chain without a single host round trip: the pose state is a float64 (B,7) tensor in HBM, every
step is a kernel of the C ABI on one stream, and the fixed-shape chain can be captured in a CUDA graph.
The result is the reference's output record per object: [qw qx qy qz tx ty tz]."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import engine, ops
from ._C import check, lib, ptr, stream


class PoseEstimator:
    def __init__(self, estimator, refiner, iterations: int = 2, precision: str = "fp32", chunk_crops: int = 256,
                 channels_last: bool = False, encoder: str = "auto"):
        # channels_last=False: with TF32 disabled cuDNN's NHWC fp32 convolutions fall back to a direct kernel that
        # is 3.5x slower than the NCHW ones (profiles/r1_call2_encoder_variants.json); NHWC only pays with TF32 on.
        self.estimator, self.refiner = estimator, refiner
        self.iterations = int(iterations)
        self.precision = precision
        self.chunk = int(chunk_crops)
        self.channels_last = channels_last
        self.n = estimator.num_points
        self.device = next(estimator.parameters()).device
        self._w_ver = None
        self._retired = []            # workspaces / packed weights a captured CUDA graph may still point to: never freed
        self.refresh()
        self._ws_head = None
        self._ws_ref = None
        self._bufs: Dict[int, dict] = {}
        # encoder: "tc" = densefusion_b200.encoder (tcgen05 implicit-GEMM convolutions, same arithmetic mode as the head),
        # "torch" = the torch/cuDNN module; "auto" = tc whenever the head runs on the tensor cores
        if encoder == "auto":
            encoder = "tc" if precision in ("3xtf32", "tf32", "hybrid", "hybrid16", "hybrid16s") else "torch"
        if encoder not in ("tc", "torch"):
            raise ValueError("encoder must be 'auto', 'tc' or 'torch'")
        self.encoder = encoder
        self._enc = None
        self.concurrent_buckets = True
        self._streams = []
        if channels_last:
            estimator.cnn.to(memory_format=torch.channels_last)

    # ---- weights ------------------------------------------------------------------------------
    def refresh(self) -> bool:
        """Re-pack the GEMM operands when the modules' parameters changed (load_state_dict, an optimiser step).  Called at the
        start of every eager estimate; CUDA graphs captured earlier keep replaying the weights they were captured with (their
        packed tensors stay alive in `_retired`) -- re-capture after a refresh that returns True."""
        ver = (engine.param_version(self.estimator), engine.param_version(self.refiner))
        if ver == self._w_ver:
            return False
        if self._w_ver is not None:
            self._retired.append((self.w_head, self.w_ref, getattr(self, "_enc", None)))
            self._enc = None
        self.w_head = engine.PackedPoseNetHead(self.estimator)
        self.w_ref = engine.PackedRefiner(self.refiner)
        self._w_ver = ver
        return True

    # ---- scratch ------------------------------------------------------------------------------
    def _workspaces(self, crops: int):
        c = min(crops, self.chunk)
        if self._ws_head is None or self._ws_head.crops < c:
            if self._ws_head is not None:          # a captured graph has the old buffers' addresses baked in: keep them allocated
                self._retired.append((self._ws_head, self._ws_ref))
            self._ws_head = engine.Workspace(c, self.n, self.device, True)
            self._ws_ref = engine.Workspace(c, self.n, self.device, False)
        return self._ws_head, self._ws_ref

    def _buffers(self, B: int) -> dict:
        b = self._bufs.get(B)
        if b is None:
            f = dict(device=self.device, dtype=torch.float32)
            b = dict(out_r=torch.empty(B, self.n, 4, **f), out_t=torch.empty(B, self.n, 3, **f),
                     out_c=torch.empty(B, self.n, 1, **f), pose=torch.empty(B, 7, device=self.device, dtype=torch.float64),
                     which=torch.empty(B, device=self.device, dtype=torch.int64),
                     new_cloud=torch.empty(B, self.n, 3, **f), r2=torch.empty(B, 4, **f), t2=torch.empty(B, 3, **f),
                     emb_pm=torch.empty(B * self.n, 32, **f),
                     # point features of every crop (head / refiner keep separate copies: the refiner's embedding
                     # branch is computed once and reused by every iteration), per-crop features and MLP scratch
                     pf_head=torch.empty(B * self.n, 384, **f), pf_ref=torch.empty(B * self.n, 384, **f),
                     g=torch.empty(B, 1024, **f), gbias=torch.empty(B, 1920, **f),
                     mlp1=torch.empty(B, 1024, **f), mlp2=torch.empty(B, 256, **f))
            self._bufs[B] = b
        return b

    def _side_streams(self, k: int):
        while len(self._streams) < k:
            self._streams.append(torch.cuda.Stream(device=self.device))
        return self._streams

    # ---- stages ---------------------------------------------------------------------------------
    def encode(self, img: torch.Tensor, choose: torch.Tensor, emb_pm_out: torch.Tensor) -> None:
        """Colour encoder + K1's gather for one (H,W) bucket; writes (b*N,32) rows."""
        if self.encoder == "tc":
            if self._enc is None:
                from .encoder import PackedEncoder
                self._enc = PackedEncoder(self.estimator.cnn)
            self._enc.forward_points(img, choose, emb_pm_out, self.precision)
            return
        if self.channels_last:
            img = img.contiguous(memory_format=torch.channels_last)
        feat = self.estimator.cnn(img)
        B, C, H, W = feat.shape
        sb, sc, sh, sw = feat.stride()
        if sh != W * sw:
            feat = feat.contiguous()
            sb, sc, sh, sw = feat.stride()
        choose = ops.i64c(choose).view(B, -1)
        check(lib.df_gather_embedding(ptr(feat), ptr(choose), ptr(emb_pm_out), None, sb, sc, sw, B, self.n, H * W,
                                      stream()), "df_gather_embedding")

    def head_and_refine(self, cloud: torch.Tensor, emb_pm: torch.Tensor, obj: torch.Tensor,
                        iterations: Optional[int] = None) -> torch.Tensor:
        """cloud (B,N,3), emb_pm (B*N,32), obj (B,) -> pose (B,7) float64.

        Per-point layers run chunk by chunk (`chunk_crops` at a time: bounded scratch, large launches); everything with one row per
        crop -- the folded global-feature bias, pose selection, the refiner's MLP towers, pose composition -- runs once
        for all B crops."""
        iters = self.iterations if iterations is None else iterations
        B, n = cloud.shape[0], self.n
        if B == 0:                                   # a frame without detections: nothing to launch
            return torch.empty(0, 7, device=self.device, dtype=torch.float64)
        if cloud.shape[1] != n:
            raise ValueError(f"PoseEstimator was built for {n} points per crop, got {cloud.shape[1]}")
        buf = self._buffers(B)
        wh, wr = self._workspaces(B)
        cloud = ops.f32c(cloud)
        obj = ops.i64c(obj).view(-1)
        s = stream()
        p = self.precision
        chunks = [(c0, min(B, c0 + self.chunk)) for c0 in range(0, B, self.chunk)]
        x_all = cloud.view(B * n, 3)
        # ---- estimate ----
        for c0, c1 in chunks:
            rows = slice(c0 * n, c1 * n)
            engine.head_features_chunk(self.w_head, wh, buf["pf_head"][rows], x_all[rows], emb_pm[rows], c1 - c0, n,
                                       buf["g"][c0:c1], p)
        engine.head_global_bias(self.w_head, buf["g"], buf["gbias"], B, p)
        for c0, c1 in chunks:
            rows = slice(c0 * n, c1 * n)
            engine.head_towers_chunk(self.w_head, wh, buf["pf_head"][rows], buf["gbias"][c0:c1], obj[c0:c1], c1 - c0, n,
                                     buf["out_r"][c0:c1], buf["out_t"][c0:c1], buf["out_c"][c0:c1], p)
        check(lib.df_select_pose(ptr(buf["out_r"]), ptr(buf["out_t"]), ptr(buf["out_c"]), ptr(x_all), B, n, ptr(buf["pose"]),
                                 ptr(buf["which"]), s), "df_select_pose")
        # ---- refine ----
        for it in range(iters):
            check(lib.df_cloud_transform(ptr(x_all), ptr(buf["pose"]), ptr(buf["new_cloud"]), B, n, s), "df_cloud_transform")
            nc = buf["new_cloud"].view(B * n, 3)
            for c0, c1 in chunks:
                rows = slice(c0 * n, c1 * n)
                engine.refiner_features_chunk(self.w_ref, wr, buf["pf_ref"][rows], nc[rows], emb_pm[rows], c1 - c0, n,
                                              buf["g"][c0:c1], p, emb_ready=it > 0)
            engine.refiner_mlp(self.w_ref, buf["g"], buf["mlp1"], buf["mlp2"], obj, B, buf["r2"], buf["t2"], p)
            check(lib.df_pose_compose(ptr(buf["pose"]), ptr(buf["r2"]), ptr(buf["t2"]), B, s), "df_pose_compose")
        return buf["pose"]

    @torch.no_grad()
    def estimate(self, img, cloud, choose, obj, iterations: Optional[int] = None) -> torch.Tensor:
        """One (H,W) bucket: img (B,3,H,W), cloud (B,N,3), choose (B,1,N), obj (B,)|(B,1) -> (B,7) f64."""
        B = cloud.shape[0]
        if not torch.cuda.is_current_stream_capturing():
            self.refresh()
        buf = self._buffers(B)
        self.encode(img, choose, buf["emb_pm"])
        return self.head_and_refine(cloud, buf["emb_pm"], obj, iterations)

    @torch.no_grad()
    def estimate_buckets(self, buckets: Sequence[dict], iterations: Optional[int] = None) -> torch.Tensor:
        """Several (H,W) buckets (dicts with img, cloud, choose, obj) -> poses (sum B,7) in bucket order."""
        buckets = [b for b in buckets if b["cloud"].shape[0] > 0]
        total = sum(b["cloud"].shape[0] for b in buckets)
        if total == 0:
            return torch.empty(0, 7, device=self.device, dtype=torch.float64)
        if not torch.cuda.is_current_stream_capturing():
            self.refresh()
        buf = self._buffers(total)
        key = ("cat", total)
        cat = self._bufs.get(key)
        if cat is None:
            cat = dict(cloud=torch.empty(total, self.n, 3, device=self.device),
                       obj=torch.empty(total, device=self.device, dtype=torch.int64))
            self._bufs[key] = cat
        # The (H,W) buckets are independent until the head: each bucket's encoder runs on its own stream (fork / join, which
        # CUDA-graph capture records as parallel branches), so the small low-resolution layers of one bucket fill the SMs
        # another bucket's tail leaves idle.
        main = torch.cuda.current_stream(self.device)
        side = self._side_streams(len(buckets))
        o = 0
        for i, b in enumerate(buckets):
            nb = b["cloud"].shape[0]
            st = side[i] if (self.concurrent_buckets and len(buckets) > 1) else main
            if st is not main:
                st.wait_stream(main)
            with torch.cuda.stream(st):
                self.encode(b["img"], b["choose"], buf["emb_pm"][o * self.n:(o + nb) * self.n])
            cat["cloud"][o:o + nb].copy_(b["cloud"])
            cat["obj"][o:o + nb].copy_(b["obj"].view(-1))
            o += nb
        if self.concurrent_buckets and len(buckets) > 1:
            for st in side[:len(buckets)]:
                main.wait_stream(st)
        return self.head_and_refine(cat["cloud"], buf["emb_pm"], cat["obj"], iterations)


class GraphedBuckets:
    """CUDA-graph capture of PoseEstimator.estimate_buckets for a fixed list of bucket shapes.
    Static device inputs are refilled from (pinned) host tensors with copy_(non_blocking=True)."""

    def __init__(self, est: PoseEstimator, shapes: Sequence[Tuple[int, int, int]], warmup: int = 2):
        # shapes: (crops, H, W) per bucket
        self.est = est
        dev, n = est.device, est.n
        self.static = [dict(img=torch.zeros(b, 3, h, w, device=dev),
                            cloud=torch.zeros(b, n, 3, device=dev),
                            choose=torch.zeros(b, 1, n, device=dev, dtype=torch.int64),
                            obj=torch.zeros(b, device=dev, dtype=torch.int64)) for b, h, w in shapes]
        for s in self.static:
            s["cloud"][..., 2] = 1.0
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                est.estimate_buckets(self.static)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = est.estimate_buckets(self.static)

    def load(self, host_buckets: Sequence[dict]) -> None:
        for s, h in zip(self.static, host_buckets):
            for k in ("img", "cloud", "choose", "obj"):
                s[k].copy_(h[k].view(s[k].shape), non_blocking=True)

    def run(self) -> torch.Tensor:
        self.graph.replay()
        return self.out


class StreamingEstimator:
    """Steady-state serving loop over fixed bucket shapes: the host->device copy of batch i+1 runs on a copy stream
    while batch i computes (two CUDA graphs over two sets of static inputs; scratch is shared because the graphs replay
    back to back on one stream), and the poses of batch i come back through a pinned buffer per slot.

        t = se.submit(host_buckets)      # pinned host tensors: img, cloud, choose, obj per bucket
        poses = se.result(t)             # (B,7) float64 pinned host tensor, valid until the slot is reused"""

    def __init__(self, est: PoseEstimator, shapes: Sequence[Tuple[int, int, int]], slots: int = 2):
        self.est = est
        dev = est.device
        self.slots = [GraphedBuckets(est, shapes) for _ in range(slots)]
        total = sum(s[0] for s in shapes)
        self.host_out = [torch.empty(total, 7, dtype=torch.float64).pin_memory() for _ in range(slots)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.loaded = [torch.cuda.Event() for _ in range(slots)]      # inputs of the slot are on the device
        self.consumed = [torch.cuda.Event() for _ in range(slots)]    # the slot's graph has read its inputs (= finished)
        self.done = [torch.cuda.Event() for _ in range(slots)]        # poses of the slot are in host memory
        self.count = 0
        main = torch.cuda.current_stream(dev)
        for e in self.consumed:
            e.record(main)

    def submit(self, host_buckets: Sequence[dict]) -> int:
        k = self.count % len(self.slots)
        main = torch.cuda.current_stream(self.est.device)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[k])             # the previous tenant of this slot is finished
            self.slots[k].load(host_buckets)
            self.loaded[k].record(self.copy_stream)
        main.wait_event(self.loaded[k])
        out = self.slots[k].run()
        self.consumed[k].record(main)
        self.host_out[k].copy_(out, non_blocking=True)
        self.done[k].record(main)
        self.count += 1
        return self.count - 1

    def result(self, ticket: int) -> torch.Tensor:
        k = ticket % len(self.slots)
        self.done[k].synchronize()
        return self.host_out[k]
