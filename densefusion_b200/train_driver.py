"""Training-driver semantics of the reference (tools/train.py:51-251) around the data-parallel step -- SURVEY.md 8f row N3.

What is kept exactly: the two-phase schedule (estimator until the best test distance drops below `refine_margin`, then
the refiner with PoseNet frozen; `batch_size //= iteration` on the switch, :92,:227), the one-time learning-rate / loss
weight decay below `decay_margin` (:219-223), the fact that every switch builds a NEW Adam (moments and step count start
from zero, :222,:228), un-normalised gradient accumulation over `batch_size` samples (:159-169), the test epoch (mean of
the per-sample distances, with the refine iterations once the refiner trains, :181-210), the checkpoint names
(`pose_model_current.pth` every 1000 samples :172-176, `pose_model_{epoch}_{dis}.pth` / `pose_refine_model_{epoch}_{dis}.pth`
on a new best :211-217) and the reference `state_dict` keys, and the resume rules (:83-97).

What changes: a "batch" is evaluated at once (crops grouped by size) instead of sample by sample, ranks split every batch
(`trainer.shard_range`) and exchange gradients once per optimiser step, and there is no host synchronisation inside a
step.  Datasets plug in through the reference's 6-tuple (points, choose, img, target, model_points, idx) and the two
accessors get_sym_list() / get_num_points_mesh(); `SyntheticPoseDataset` provides that contract without files."""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import synth
from .trainer import DataParallelTrainer, shard_range


@dataclass
class TrainOptions:                      # defaults of tools/train.py:31-48
    batch_size: int = 8
    lr: float = 0.0001
    lr_rate: float = 0.3
    w: float = 0.015
    w_rate: float = 0.3
    decay_margin: float = 0.016
    refine_margin: float = 0.013
    iteration: int = 2
    nepoch: int = 500
    repeat_epoch: int = 1
    start_epoch: int = 1
    outf: str = "trained_models"
    resume_posenet: str = ""
    resume_refinenet: str = ""
    save_every: int = 1000


class Schedule:
    """The phase / decay state machine of tools/train.py (pure host logic)."""

    def __init__(self, opt: TrainOptions):
        self.opt = opt
        self.lr, self.w, self.batch_size = opt.lr, opt.w, opt.batch_size
        self.decay_start = self.refine_start = False
        self.best_test = float("inf")
        if opt.resume_refinenet:                     # tools/train.py:86-93
            self.refine_start = self.decay_start = True
            self.lr *= opt.lr_rate
            self.w *= opt.w_rate
            self.batch_size = int(self.batch_size / opt.iteration)

    @property
    def phase(self) -> str:
        return "refiner" if self.refine_start else "estimator"

    def checkpoint_name(self, epoch: int, test_dis: float) -> str:
        stem = "pose_refine_model" if self.refine_start else "pose_model"
        return "{0}_{1}_{2}.pth".format(stem, epoch, test_dis)

    def current_name(self) -> str:
        return "pose_refine_model_current.pth" if self.refine_start else "pose_model_current.pth"

    def after_test(self, epoch: int, test_dis: float) -> List[Tuple[str, object]]:
        """tools/train.py:211-251.  Returns the actions in the reference's order:
        ('save_best', name) / ('decay', (lr, w)) / ('refine_start', batch_size)."""
        actions: List[Tuple[str, object]] = []
        if test_dis <= self.best_test:
            self.best_test = test_dis
            actions.append(("save_best", self.checkpoint_name(epoch, test_dis)))
        if self.best_test < self.opt.decay_margin and not self.decay_start:
            self.decay_start = True
            self.lr *= self.opt.lr_rate
            self.w *= self.opt.w_rate
            actions.append(("decay", (self.lr, self.w)))
        if self.best_test < self.opt.refine_margin and not self.refine_start:
            self.refine_start = True
            self.batch_size = int(self.batch_size / self.opt.iteration)
            actions.append(("refine_start", self.batch_size))
        return actions


class SyntheticPoseDataset:
    """The dataset contract of datasets/ycb/dataset.py:97-232 without files: __getitem__ -> (points (N,3), choose (1,N),
    img (3,H,W), target (M,3), model_points (M,3), idx (1,)); M = 500, or 2600 once the refiner trains (:90-91,:205-208)."""

    def __init__(self, mode: str, num_points: int, length: int, num_obj: int = 21, refine: bool = False, seed: int = 0,
                 sizes: Sequence[Tuple[int, int]] = ((80, 80), (120, 120), (160, 160))):
        self.mode, self.n, self.length, self.o, self.refine, self.seed, self.sizes = mode, num_points, length, num_obj, refine, seed, sizes

    def __len__(self):
        return self.length

    def get_sym_list(self):
        return list(synth.YCB_SYM)

    def get_num_points_mesh(self):
        return 2600 if self.refine else 500

    def __getitem__(self, i):
        case = self.seed * 100003 + (0 if self.mode == "train" else 50000) + i
        d = synth.synth_crop(case, self.n, self.get_num_points_mesh(), self.o, self.sizes[i % len(self.sizes)])
        return d["points"][0], d["choose"][0], d["img"][0], d["target"][0], d["model_points"][0], d["idx"][0]


def collate_buckets(samples, device) -> List[dict]:
    """Stack 6-tuples of equal crop size into the batched dicts DataParallelTrainer.step expects."""
    groups: Dict[Tuple[int, int], list] = {}
    for s in samples:
        groups.setdefault(tuple(s[2].shape[1:]), []).append(s)
    out = []
    for _, g in sorted(groups.items()):
        out.append({"points": torch.stack([s[0] for s in g]).to(device), "choose": torch.stack([s[1] for s in g]).to(device),
                    "img": torch.stack([s[2] for s in g]).to(device), "target": torch.stack([s[3] for s in g]).to(device),
                    "model_points": torch.stack([s[4] for s in g]).to(device), "idx": torch.stack([s[5] for s in g]).to(device)})
    return out


class TrainDriver:
    def __init__(self, estimator, refiner, dataset_factory: Callable[[str, bool], object], opt: TrainOptions,
                 rank: int = 0, world: int = 1, log: Callable[[str], None] = print):
        self.est, self.ref, self.factory, self.opt, self.rank, self.world, self.log = estimator, refiner, dataset_factory, opt, rank, world, log
        self.device = next(estimator.parameters()).device
        self.sched = Schedule(opt)
        if opt.resume_posenet:
            estimator.load_state_dict(torch.load(os.path.join(opt.outf, opt.resume_posenet), map_location=self.device))
        if opt.resume_refinenet:
            refiner.load_state_dict(torch.load(os.path.join(opt.outf, opt.resume_refinenet), map_location=self.device))
        self._load_data()
        self.trainer = DataParallelTrainer(estimator, refiner, self.num_points_mesh, self.sym_list, lr=self.sched.lr,
                                           w=self.sched.w, iteration=opt.iteration, phase=self.sched.phase)
        self.samples_seen = 0

    def _load_data(self):
        self.train_set = self.factory("train", self.sched.refine_start)
        self.test_set = self.factory("test", self.sched.refine_start)
        self.sym_list = self.train_set.get_sym_list()
        self.num_points_mesh = self.train_set.get_num_points_mesh()

    def _save(self, name: str):
        if self.rank == 0:
            os.makedirs(self.opt.outf, exist_ok=True)
            net = self.ref if name.startswith("pose_refine_model") else self.est      # (the schedule may already have switched)
            torch.save(net.state_dict(), os.path.join(self.opt.outf, name))

    # ---- one training epoch: optimiser steps of world * batch_size samples ----
    def train_epoch(self, epoch: int, order: Optional[Sequence[int]] = None) -> float:
        bs = self.sched.batch_size
        self.est.eval() if self.sched.refine_start else self.est.train()
        self.ref.train()
        order = list(range(len(self.train_set))) if order is None else list(order)
        dis_sum, count = 0.0, 0
        for _ in range(self.opt.repeat_epoch):
            for start in range(0, len(order) - bs * self.world + 1, bs * self.world):
                group = order[start:start + bs * self.world]
                mine = [group[i] for i in shard_range(len(group), self.rank, self.world)]
                out = self.trainer.step(collate_buckets([self.train_set[i] for i in mine], self.device))
                dis_sum += float(out["dis_sum"])
                count += len(mine)
                before = self.samples_seen
                self.samples_seen += len(group)
                if before // self.opt.save_every != self.samples_seen // self.opt.save_every:      # tools/train.py:172-176
                    self._save(self.sched.current_name())
        return dis_sum / max(count, 1)

    # ---- test epoch (tools/train.py:181-210): mean per-sample distance, refined once the refiner trains ----
    @torch.no_grad()
    def test_epoch(self) -> float:
        self.est.eval()
        self.ref.eval()
        mine = list(shard_range(len(self.test_set), self.rank, self.world))
        total = torch.zeros(2, device=self.device, dtype=torch.float64)
        tr = self.trainer
        for start in range(0, len(mine), 32):
            for b in collate_buckets([self.test_set[i] for i in mine[start:start + 32]], self.device):
                r, t, c, emb = self.est.forward_batched(b["img"], b["points"], b["choose"], b["idx"])
                _, dis, pts, tgt = tr.criterion(r, t, c, b["target"], b["model_points"], b["idx"], b["points"], self.sched.w,
                                                self.sched.refine_start)
                if self.sched.refine_start:
                    for _ in range(self.opt.iteration):
                        pr, pt = self.ref.forward_batched(pts, emb, b["idx"])
                        dis, pts, tgt = tr.criterion_refine(pr, pt, tgt, b["model_points"], b["idx"], pts)
                total[0] += dis.double().sum()
                total[1] += dis.numel()
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(total)
        return float(total[0] / total[1].clamp(min=1))

    def _apply(self, actions):
        for kind, arg in actions:
            if kind == "save_best":
                self._save(arg)
            elif kind == "decay":                     # new Adam for the estimator with the decayed rate (tools/train.py:219-223)
                self.trainer.lr, self.trainer.w = arg
                self.trainer.reset_optimizer()
            elif kind == "refine_start":              # new Adam for the refiner; datasets / losses rebuilt (tools/train.py:225-251)
                self._load_data()
                from .lib.loss import Loss
                from .lib.loss_refiner import Loss_refine
                self.trainer.criterion = Loss(self.num_points_mesh, self.sym_list)
                self.trainer.criterion_refine = Loss_refine(self.num_points_mesh, self.sym_list)
                self.trainer.set_phase("refiner")
                self.trainer.reset_optimizer()

    def run(self, epochs: Optional[int] = None) -> List[dict]:
        history = []
        last = self.opt.nepoch if epochs is None else self.opt.start_epoch + epochs
        for epoch in range(self.opt.start_epoch, last):
            train_dis = self.train_epoch(epoch)
            test_dis = self.test_epoch()
            actions = self.sched.after_test(epoch, test_dis)
            self._apply(actions)
            history.append({"epoch": epoch, "train_dis": train_dis, "test_dis": test_dis, "phase": self.sched.phase,
                            "lr": self.sched.lr, "w": self.sched.w, "batch_size": self.sched.batch_size,
                            "actions": [a[0] for a in actions]})
            self.log(f"epoch {epoch}: train dis {train_dis:.6f} test dis {test_dis:.6f} -> {history[-1]['actions']}")
        return history
