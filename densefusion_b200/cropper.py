"""Device-side input preparation: raw frames -> the tensors PoseEstimator consumes.

Reference: tools/eval_ycb.py:54-91 (get_bbox: PoseCNN box snapped to multiples of 40 px and clamped into the image) and
:147-190 (mask, `choose`, back-projected cloud, normalised colour crop), done there per object in numpy followed by four
host->device copies.  Here the host only derives the integer boxes; everything per pixel runs in df_build_crops, one
launch pair per crop-size bucket, and the result is already grouped the way pipeline.PoseEstimator.estimate_buckets wants."""
from __future__ import annotations

import ctypes
from typing import Dict, List, Sequence, Tuple

import torch

from ._C import check, lib, ptr, stream

BORDER = tuple(range(40, 681, 40))              # tools/eval_ycb.py:34 (border_list without its -1 sentinel)
YCB_CAMERA = (312.9869, 241.3109, 1066.778, 1067.487, 10000.0)          # cx, cy, fx, fy, depth scale (eval_ycb.py:37-41)
IMAGENET_MEAN_STD = (0.485, 0.456, 0.406, 0.229, 0.224, 0.225)          # transforms.Normalize of eval_ycb.py:33


def _snap(extent: int) -> int:
    """Next multiple of 40 strictly above `extent`, unless it already is one (the reference's strict comparisons leave
    exact multiples -- and anything beyond 680 -- unchanged)."""
    prev = -1
    for b in BORDER:
        if prev < extent < b:
            return b
        prev = b
    return extent


def get_bbox(roi: Sequence[float], img_h: int = 480, img_w: int = 640) -> Tuple[int, int, int, int]:
    """roi = one PoseCNN row [batch, class, x1, y1, x2, y2]; returns (rmin, rmax, cmin, cmax) like eval_ycb.py:54-91."""
    rmin, rmax = int(roi[3]) + 1, int(roi[5]) - 1
    cmin, cmax = int(roi[2]) + 1, int(roi[4]) - 1
    r_b, c_b = _snap(rmax - rmin), _snap(cmax - cmin)
    cr, cc = int((rmin + rmax) / 2), int((cmin + cmax) / 2)
    rmin, rmax = cr - int(r_b / 2), cr + int(r_b / 2)
    cmin, cmax = cc - int(c_b / 2), cc + int(c_b / 2)
    if rmin < 0:
        rmin, rmax = 0, rmax - rmin
    if cmin < 0:
        cmin, cmax = 0, cmax - cmin
    if rmax > img_h:
        rmin, rmax = rmin - (rmax - img_h), img_h
    if cmax > img_w:
        cmin, cmax = cmin - (cmax - img_w), img_w
    return rmin, rmax, cmin, cmax


class CropBuilder:
    def __init__(self, num_points: int, camera: Sequence[float] = YCB_CAMERA, mean_std: Sequence[float] = IMAGENET_MEAN_STD):
        self.n = int(num_points)
        self._cam = (ctypes.c_float * 5)(*camera)
        self._ms = (ctypes.c_float * 6)(*mean_std)

    def build(self, rgb: torch.Tensor, depth: torch.Tensor, label: torch.Tensor,
              objects: Sequence[Tuple[int, int, Sequence[float]]], seed: int = 0) -> List[dict]:
        """rgb (F,H,W,3) uint8, depth (F,H,W) fp32, label (F,H,W) int32 CUDA tensors; objects = (frame, item id, roi row).
        Returns one dict per crop size: img (b,3,h,w), cloud (b,N,3), choose (b,1,N), obj (b,) = item id - 1
        (eval_ycb.py:180), count (b,) masked pixels (0 = lost object) and `order` (positions in `objects`)."""
        F, H, W, _ = rgb.shape
        groups: Dict[Tuple[int, int], list] = {}
        for pos, (frame, item, roi) in enumerate(objects):
            rmin, rmax, cmin, cmax = get_bbox(roi, H, W)
            groups.setdefault((rmax - rmin, cmax - cmin), []).append((pos, frame, item, rmin, rmax, cmin, cmax))
        out = []
        dev = rgb.device
        for (h, w), rows in sorted(groups.items()):
            b = len(rows)
            meta = torch.tensor([r[1:] for r in rows], dtype=torch.int32).to(dev)
            img = torch.empty(b, 3, h, w, device=dev, dtype=torch.float32)
            choose = torch.empty(b, self.n, device=dev, dtype=torch.int64)
            cloud = torch.empty(b, self.n, 3, device=dev, dtype=torch.float32)
            count = torch.empty(b, device=dev, dtype=torch.int32)
            check(lib.df_build_crops(ptr(rgb), ptr(depth), ptr(label), ptr(meta), b, H, W, h, w, self.n, self._cam, self._ms,
                                     int(seed) & 0xffffffff, ptr(img), ptr(choose), ptr(cloud), ptr(count), stream()),
                  "df_build_crops")
            out.append({"img": img, "cloud": cloud, "choose": choose.view(b, 1, self.n),
                        "obj": torch.tensor([r[2] - 1 for r in rows], dtype=torch.int64, device=dev), "count": count,
                        "order": [r[0] for r in rows]})
        return out
