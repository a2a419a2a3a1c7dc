"""Seeded synthetic inputs and weights.

Everything is generated on the CPU with explicit torch.Generator objects so that the oracle, the
golden-fixture script (which runs the reference in the build container) and the CUDA path all see
identical bits, on any machine with the same torch build.  There is no network, hence no dataset
or checkpoint: shapes follow datasets/ycb/dataset.py:97-232 and datasets/linemod/dataset.py:90-195.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Iterable, Tuple

import torch

YCB_SYM = [12, 15, 18, 19, 20]      # datasets/ycb/dataset.py:89
LINEMOD_SYM = [7, 8]                # datasets/linemod/dataset.py:88


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed) & 0x7FFFFFFFFFFFFFFF)
    return g


def synth_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic weights keyed BY PARAMETER NAME (independent of module construction order, so
    the same values can be loaded into the reference modules and into the drop-ins).

    Conv2d of the encoder: N(0, sqrt(2/(k*k*out))) like lib/extractors.py:91-97.  Conv1d / Linear /
    other weights: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (torch's default bound).  PReLU slope 0.25."""
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        g = _gen(seed * 1000003 + zlib.crc32(name.encode()))
        if len(shape) == 1 and shape[0] == 1 and name.endswith(".weight"):        # PReLU
            w = torch.full(shape, 0.25)
        elif len(shape) == 4 and name.endswith(".weight"):
            n = shape[2] * shape[3] * shape[0]
            w = torch.randn(shape, generator=g) * math.sqrt(2.0 / n)
        elif name.endswith(".weight"):
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            b = 1.0 / math.sqrt(max(fan_in, 1))
            w = (torch.rand(shape, generator=g) * 2.0 - 1.0) * b
        else:                                                                      # bias
            w = (torch.rand(shape, generator=g) * 2.0 - 1.0) * 0.05
        out[name] = w.float().contiguous()
    return out


def shapes_of(module: torch.nn.Module) -> Dict[str, Tuple[int, ...]]:
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}


def random_unit_quaternion(g: torch.Generator) -> torch.Tensor:
    q = torch.randn(4, generator=g)
    return q / q.norm()


def quat_to_rot(q: torch.Tensor) -> torch.Tensor:
    w, x, y, z = [float(v) for v in q]
    return torch.tensor([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ], dtype=torch.float32)


def synth_crop(case: int, num_points: int = 500, num_pt_mesh: int = 500, num_obj: int = 21,
               hw: Tuple[int, int] = (80, 80), obj: int | None = None):
    """One object crop in the reference's 6-tuple layout (+ image), batch dim 1:
    img (1,3,H,W) f32, points (1,N,3) f32, choose (1,1,N) i64, target (1,M,3), model_points (1,M,3),
    idx (1,1) i64."""
    g = _gen(1234 + case)
    h, w = hw
    img = torch.randn(1, 3, h, w, generator=g)
    perm = torch.randperm(h * w, generator=g)
    if h * w >= num_points:
        choose = torch.sort(perm[:num_points])[0]
    else:                                                   # wrap-pad like the dataset does
        reps = (num_points + h * w - 1) // (h * w)
        choose = torch.sort(perm.repeat(reps)[:num_points])[0]
    choose = choose.view(1, 1, num_points).long()
    points = torch.randn(1, num_points, 3, generator=g) * 0.05 + torch.tensor([0.0, 0.0, 0.8])
    model_points = torch.randn(1, num_pt_mesh, 3, generator=g) * 0.05
    rot = quat_to_rot(random_unit_quaternion(g))
    t_gt = torch.tensor([0.0, 0.0, 0.8])
    target = model_points @ rot.t() + t_gt
    if obj is None:
        obj = int(torch.randint(0, num_obj, (1,), generator=g).item())
    idx = torch.tensor([[obj]], dtype=torch.int64)
    return {"img": img.contiguous(), "points": points.contiguous(), "choose": choose.contiguous(),
            "target": target.contiguous(), "model_points": model_points.contiguous(), "idx": idx}


def synth_embedding(case: int, num_points: int = 500) -> torch.Tensor:
    """Head-only runs: emb (1,32,N) = log_softmax(N(0,1)) over channels, the encoder's output law
    (lib/pspnet.py:53-56)."""
    g = _gen(77000 + case)
    return torch.log_softmax(torch.randn(1, 32, num_points, generator=g), dim=1).contiguous()


def synth_predictions(case: int, num_points: int = 500):
    """Head-independent (pred_r, pred_t, pred_c) for loss tests: un-normalised quaternions,
    small offsets, confidences in (0,1) with one clear maximum plus exact ties elsewhere."""
    g = _gen(555000 + case)
    pred_r = torch.randn(1, num_points, 4, generator=g)
    pred_t = torch.randn(1, num_points, 3, generator=g) * 0.02
    pred_c = torch.rand(1, num_points, 1, generator=g) * 0.9 + 0.05
    return pred_r.contiguous(), pred_t.contiguous(), pred_c.contiguous()


def batch_crops(cases: Iterable[int], **kw):
    """Stack crops of equal (H,W) into batched tensors (B,...) for the batched entry points."""
    items = [synth_crop(c, **kw) for c in cases]
    return {k: torch.cat([it[k] for it in items], 0).contiguous() for k in items[0]}
