"""Result record and accuracy metrics of the evaluation drivers.

  * pose record: the reference stores, per frame, an (n,7) array [qw qx qy qz tx ty tz] under the key 'poses' of a
    MATLAB file (tools/eval_ycb.py:233-240; a lost detection is a row of zeros, :234-237) -- `save_poses_mat`.
  * per-object pose error (tools/eval_linemod.py:118-131): model points under the estimated pose against the ground
    truth points; symmetric objects are matched through the 1-NN first.  This is exactly what the fused K3 kernel
    computes for a single hypothesis, so `pose_distances` is one df_loss_forward launch for a whole batch of poses.
  * YCB-Video toolbox metrics (replace_ycb_toolbox/evaluate_poses_keyframe.m:160-193 add / adi,
    plot_accuracy_keyframe.m:37-53,150-165 AUC): `adi_distances` (nearest ESTIMATED point for every ground-truth point,
    the toolbox's direction) on df_knn, `auc` = VOCap restated in numpy.
The metric helpers around the kernels are small host-side array code; nothing here is on the pose hot path."""
from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np
import torch

from . import ops


def save_poses_mat(path: str, poses) -> None:
    """poses: (n,7) array-like [qw qx qy qz tx ty tz] (float64), zero rows for objects that were not detected."""
    import scipy.io as scio
    arr = np.asarray(poses.detach().cpu() if torch.is_tensor(poses) else poses, dtype=np.float64).reshape(-1, 7)
    scio.savemat(path, {"poses": arr})


def load_poses_mat(path: str) -> np.ndarray:
    import scipy.io as scio
    return np.asarray(scio.loadmat(path)["poses"], dtype=np.float64).reshape(-1, 7)


def pose_distances(poses: torch.Tensor, model_points: torch.Tensor, target: torch.Tensor, idx: torch.Tensor,
                   sym_list: Iterable[int]) -> torch.Tensor:
    """tools/eval_linemod.py:118-131 for a batch: poses (B,7) [q t], model_points / target (B,M,3), idx (B,) ->
    dis (B,) fp32: mean_j |R m_j + t - target_j|, with target re-indexed by the nearest neighbour of every
    transformed model point for objects in sym_list (ADD-S), on the GPU (df_loss_forward, single hypothesis)."""
    B = poses.shape[0]
    q = poses[:, :4].to(torch.float32).contiguous().view(B, 1, 4)
    t = poses[:, 4:].to(torch.float32).contiguous().view(B, 1, 3)
    dummy = torch.zeros(B, 1, 3, device=poses.device, dtype=torch.float32)
    st = ops.loss_forward(q, t, None, target, model_points, None, dummy, idx, ops.sym_mask(sym_list), True, 0.0)
    return st.dis_sel


def success_rates(dis, idx, diameters: Sequence[float], num_objects: int):
    """tools/eval_linemod.py:133-145: per-object and overall share of poses with dis < diameter[obj]
    (the caller passes 0.1 * object diameter, tools/eval_linemod.py:57-61)."""
    dis = np.asarray(dis.detach().cpu() if torch.is_tensor(dis) else dis, dtype=np.float64).reshape(-1)
    idx = np.asarray(idx.detach().cpu() if torch.is_tensor(idx) else idx).reshape(-1).astype(np.int64)
    ok = np.zeros(num_objects, dtype=np.int64)
    cnt = np.zeros(num_objects, dtype=np.int64)
    for d, o in zip(dis, idx):
        cnt[o] += 1
        ok[o] += 1 if d < diameters[o] else 0
    per_obj = np.where(cnt > 0, ok / np.maximum(cnt, 1), np.nan)
    return per_obj, float(ok.sum()) / max(int(cnt.sum()), 1)


def _rotation(q: torch.Tensor) -> torch.Tensor:
    """(B,4) quaternions (w,x,y,z) -> (B,3,3), normalised like lib/transformations.py:1254-1278."""
    q = q.double()
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1).view(-1, 3, 3)


def add_distances(poses, poses_gt, points) -> torch.Tensor:
    """evaluate_poses_keyframe.m:160-174 `add`: mean distance between the model points under both poses.  points (M,3)."""
    pe = points.double() @ _rotation(poses[:, :4]).transpose(1, 2) + poses[:, None, 4:].double()
    pg = points.double() @ _rotation(poses_gt[:, :4]).transpose(1, 2) + poses_gt[:, None, 4:].double()
    return (pe - pg).norm(dim=2).mean(dim=1)


def adi_distances(poses, poses_gt, points) -> torch.Tensor:
    """evaluate_poses_keyframe.m:176-193 `adi`: for every ground-truth point the distance to the NEAREST estimated point
    (the toolbox's direction; KDTreeSearcher there, the bit-exact df_knn kernel here)."""
    pe = (points.double() @ _rotation(poses[:, :4]).transpose(1, 2) + poses[:, None, 4:].double()).float()
    pg = (points.double() @ _rotation(poses_gt[:, :4]).transpose(1, 2) + poses_gt[:, None, 4:].double()).float()
    inds = ops.knn(pe.transpose(1, 2).contiguous(), pg.transpose(1, 2).contiguous(), 1).squeeze(1) - 1      # (B,M)
    near = torch.gather(pe, 1, inds.unsqueeze(-1).expand(-1, -1, 3))
    return (pg - near).norm(dim=2).mean(dim=1).double()


def auc(distances, max_distance: float = 0.1) -> float:
    """plot_accuracy_keyframe.m:41-53 + VOCap (:150-165): area under the accuracy-vs-threshold curve up to 10 cm, in [0,1]."""
    d = np.sort(np.asarray(distances.detach().cpu() if torch.is_tensor(distances) else distances, dtype=np.float64).reshape(-1))
    d = np.where(d > max_distance, np.inf, d)
    n = d.size
    if n == 0:
        return 0.0
    acc = np.cumsum(np.ones(n)) / n
    keep = np.isfinite(d)
    rec, prec = d[keep], acc[keep]
    if rec.size == 0:
        return 0.0
    mrec = np.concatenate([[0.0], rec, [max_distance]])
    mpre = np.concatenate([[0.0], prec, [prec[-1]]])
    mpre = np.maximum.accumulate(mpre)
    i = np.nonzero(mrec[1:] != mrec[:-1])[0] + 1
    return float(np.sum((mrec[i] - mrec[i - 1]) * mpre[i]) * 10.0)
