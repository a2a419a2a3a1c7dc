"""oracle/df_oracle.py -- CPU restatement of the reference's per-pixel pose-hypothesis path.

TEST INFRASTRUCTURE ONLY.  Nothing under densefusion_b200/ imports this file; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may, and there
only as the checker (or the timed CPU baseline), never as the product.

Parity status: the reference's own tests hold NO golden vector for the network / loss /
selection / kNN part of the path, so for those rows the oracle is pinned
against outputs of the reference itself, generated in the build container by importing
/root/reference (tests/golden/make_golden.py -> tests/golden/*.npz, checked by
tests/test_oracle_golden.py).  The two quaternion helpers ARE pinned by the reference's own
known-answer doctests (lib/transformations.py:1257-1265, :1287-1317; tests/test_oracle_golden.py).

Everything here is fp32 torch-on-CPU (autograd enabled, so it is also the gradient oracle),
except the host-side pose algebra which is float64 numpy exactly like the reference's eval loop.
Weights are taken from a reference-layout state_dict (same keys the reference modules use).

Restated functions (reference file:line):
  psp_encoder            lib/network.py:27-37, lib/pspnet.py:7-77, lib/extractors.py:20-124
  gather_embedding       lib/network.py:98-102
  posenet_feat           lib/network.py:53-68
  posenet_head           lib/network.py:104-130
  posenet_forward        lib/network.py:95-132
  refiner_forward        lib/network.py:151-168, :187-206
  loss                   lib/loss.py:13-70      (ADD-S branch with the upstream kNN contract,
                                                 lib/knn/__init__.py:15-23)
  loss_refine            lib/loss_refiner.py:12-62
  quaternion_matrix      lib/transformations.py:1254-1278
  quaternion_from_matrix lib/transformations.py:1281-1363 (isprecise branch + sign fix)
  select_pose / refine_pose_eval   tools/eval_ycb.py:193-233
  knn                    lib/knn/src/knn_cuda_kernel.cu:31-170 via oracle/knn_ref.c
  estimator_gradients / refiner_gradients / adam_reference   tools/train.py:143-169 (+ torch.optim.Adam)
  crop_bbox / build_crop tools/eval_ycb.py:54-91, :147-178 (numpy, like the reference)
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libdf_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle C library missing: run `make -C oracle` (or __graft_entry__.build())")
        lib = ctypes.CDLL(path)
        lib.df_oracle_knn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        lib.df_oracle_knn.restype = ctypes.c_int
        lib.df_oracle_knn1_d3.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_void_p]
        lib.df_oracle_knn1_d3.restype = ctypes.c_int
        lib.df_oracle_num_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def num_threads() -> int:
    return int(_lib().df_oracle_num_threads())


# ----------------------------------------------------------------------------------------------
# kNN (lib/knn) -- bit-exact emulator
# ----------------------------------------------------------------------------------------------
def knn(ref: torch.Tensor, query: torch.Tensor, k: int = 1) -> torch.Tensor:
    """ref (B,D,R) f32, query (B,D,Q) f32 -> int64 (B,k,Q), 1-based (lib/knn/__init__.py:15-23)."""
    assert ref.dim() == 3 and query.dim() == 3 and ref.shape[:2] == query.shape[:2]
    ref = ref.detach().float().contiguous().cpu()
    query = query.detach().float().contiguous().cpu()
    B, D, R = ref.shape
    Q = query.shape[2]
    out = torch.empty(B, k, Q, dtype=torch.int64)
    lib = _lib()
    for b in range(B):
        if D == 3 and k == 1:
            rc = lib.df_oracle_knn1_d3(ref[b].data_ptr(), R, query[b].data_ptr(), Q, out[b].data_ptr())
        else:
            rc = lib.df_oracle_knn(ref[b].data_ptr(), R, query[b].data_ptr(), Q, D, k, out[b].data_ptr())
        if rc != 0:
            raise RuntimeError("df_oracle_knn: bad arguments")
    return out


def knn_numpy_small(ref: np.ndarray, query: np.ndarray) -> np.ndarray:
    """Pure-numpy 1-NN for tiny cases (cross-check of the C emulator). ref (D,R), query (D,Q)."""
    D, R = ref.shape
    Q = query.shape[1]
    out = np.empty(Q, dtype=np.int64)
    for q in range(Q):
        best, arg = None, 1
        for r in range(R):
            ssd = np.float32(0.0)
            for d in range(D):
                tmp = np.float32(ref[d, r]) - np.float32(query[d, q])
                # fmaf: exact product+sum in float64 is exact for fp32 inputs only when it fits in
                # 53 bits; use Python's math.fma-free emulation through float64 (24+24 bit product
                # is exact in 53 bits; the add may round, so round once via longdouble).
                ssd = np.float32(np.longdouble(tmp) * np.longdouble(tmp) + np.longdouble(ssd))
            if best is None:
                best = ssd
            elif ssd < best:
                best, arg = ssd, r + 1
        out[q] = arg
    return out


# ----------------------------------------------------------------------------------------------
# CNN encoder (kept as torch/cuDNN in the product; restated functionally here)
# ----------------------------------------------------------------------------------------------
def _basic_block(sd, prefix, x, stride, dilation, has_down):
    y = F.conv2d(x, sd[prefix + "conv1.weight"], None, stride=stride, padding=dilation, dilation=dilation)
    y = F.relu(y)
    y = F.conv2d(y, sd[prefix + "conv2.weight"], None, stride=1, padding=dilation, dilation=dilation)
    res = F.conv2d(x, sd[prefix + "downsample.0.weight"], None, stride=stride) if has_down else x
    return F.relu(y + res)


def psp_encoder(sd: dict, img: torch.Tensor, prefix: str = "cnn.model.module.") -> torch.Tensor:
    """img (bs,3,H,W) -> (bs,32,H,W) per-pixel log-softmax embedding (eval mode: dropout off)."""
    p = prefix + "feats."
    x = F.relu(F.conv2d(img, sd[p + "conv1.weight"], None, stride=2, padding=3))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    # (stride of first block, dilation of the *non-first* blocks) -- extractors.py:99-112: the first
    # block of each stage is built with dilation=1 (the argument is not forwarded), the rest get it.
    cfg = [("layer1", 1, 1, False), ("layer2", 2, 1, True), ("layer3", 1, 2, True), ("layer4", 1, 4, True)]
    for name, stride, dil, down in cfg:
        x = _basic_block(sd, f"{p}{name}.0.", x, stride, 1, down)
        x = _basic_block(sd, f"{p}{name}.1.", x, 1, dil, False)
    f = x
    p = prefix + "psp."
    h, w = f.shape[2], f.shape[3]
    priors = []
    for i, size in enumerate((1, 2, 3, 6)):
        s = F.adaptive_avg_pool2d(f, (size, size))
        s = F.conv2d(s, sd[f"{p}stages.{i}.1.weight"], None)
        priors.append(F.interpolate(s, size=(h, w), mode="bilinear", align_corners=False))
    priors.append(f)
    x = F.relu(F.conv2d(torch.cat(priors, 1), sd[p + "bottleneck.weight"], sd[p + "bottleneck.bias"]))
    for up in ("up_1", "up_2", "up_3"):
        q = f"{prefix}{up}.conv."
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
        x = F.conv2d(x, sd[q + "1.weight"], sd[q + "1.bias"], padding=1)
        x = F.prelu(x, sd[q + "2.weight"])
    x = F.conv2d(x, sd[prefix + "final.0.weight"], sd[prefix + "final.0.bias"])
    return F.log_softmax(x, dim=1)


def gather_embedding(out_img: torch.Tensor, choose: torch.Tensor) -> torch.Tensor:
    """out_img (bs,32,H,W), choose (bs,1,N) i64 -> emb (bs,32,N)."""
    bs, di = out_img.shape[:2]
    flat = out_img.reshape(bs, di, -1)
    return torch.gather(flat, 2, choose.expand(bs, di, choose.shape[2])).contiguous()


# ----------------------------------------------------------------------------------------------
# Dense-fusion head
# ----------------------------------------------------------------------------------------------
def _pw(sd, name, x):
    """1x1 Conv1d on (bs,C,N) with weight (O,C,1)."""
    return F.conv1d(x, sd[name + ".weight"], sd[name + ".bias"])


def posenet_feat(sd: dict, x: torch.Tensor, emb: torch.Tensor, prefix: str = "feat.") -> torch.Tensor:
    """x (bs,3,N), emb (bs,32,N) -> (bs,1408,N): [x1 64 | e1 64 | x2 128 | e2 128 | global 1024]."""
    n = x.shape[2]
    x1 = F.relu(_pw(sd, prefix + "conv1", x))
    e1 = F.relu(_pw(sd, prefix + "e_conv1", emb))
    x2 = F.relu(_pw(sd, prefix + "conv2", x1))
    e2 = F.relu(_pw(sd, prefix + "e_conv2", e1))
    pf2 = torch.cat([x2, e2], 1)
    h = F.relu(_pw(sd, prefix + "conv5", pf2))
    h = F.relu(_pw(sd, prefix + "conv6", h))
    g = F.avg_pool1d(h, n)                                   # (bs,1024,1)
    return torch.cat([x1, e1, pf2, g.expand(-1, -1, n)], 1)


def posenet_head(sd: dict, x: torch.Tensor, emb: torch.Tensor, obj: torch.Tensor, num_obj: int):
    """x (bs,N,3) cloud, emb (bs,32,N), obj (bs,1) -> out_rx (1,N,4), out_tx (1,N,3), out_cx (1,N,1).
    Only batch element 0 is returned, as in lib/network.py:123-126."""
    bs, n = x.shape[0], x.shape[1]
    ap = posenet_feat(sd, x.transpose(2, 1).contiguous(), emb)
    outs = []
    for br, width in (("r", 4), ("t", 3), ("c", 1)):
        h = ap
        for layer in (1, 2, 3):
            h = F.relu(_pw(sd, f"conv{layer}_{br}", h))
        h = _pw(sd, f"conv4_{br}", h)
        if br == "c":
            h = torch.sigmoid(h)
        h = h.view(bs, num_obj, width, n)
        sel = torch.index_select(h[0], 0, obj[0])             # (1,width,N)
        outs.append(sel.transpose(2, 1).contiguous())
    return outs[0], outs[1], outs[2]


def posenet_forward(sd: dict, img, x, choose, obj, num_obj: int):
    out_img = psp_encoder(sd, img)
    emb = gather_embedding(out_img, choose)
    r, t, c = posenet_head(sd, x, emb, obj, num_obj)
    return r, t, c, emb.detach()


def refiner_feat(sd: dict, x: torch.Tensor, emb: torch.Tensor, prefix: str = "feat.") -> torch.Tensor:
    n = x.shape[2]
    x1 = F.relu(_pw(sd, prefix + "conv1", x))
    e1 = F.relu(_pw(sd, prefix + "e_conv1", emb))
    x2 = F.relu(_pw(sd, prefix + "conv2", x1))
    e2 = F.relu(_pw(sd, prefix + "e_conv2", e1))
    pf3 = torch.cat([x1, e1, x2, e2], 1)
    h = F.relu(_pw(sd, prefix + "conv5", pf3))
    h = F.relu(_pw(sd, prefix + "conv6", h))
    return F.avg_pool1d(h, n).view(-1, 1024)


def refiner_forward(sd: dict, x: torch.Tensor, emb: torch.Tensor, obj: torch.Tensor, num_obj: int):
    """x (bs,N,3), emb (bs,32,N), obj (bs,1) -> out_rx (1,4), out_tx (1,3) (batch element 0)."""
    bs = x.shape[0]
    g = refiner_feat(sd, x.transpose(2, 1).contiguous(), emb)
    outs = []
    for br, width in (("r", 4), ("t", 3)):
        h = F.relu(F.linear(g, sd[f"conv1_{br}.weight"], sd[f"conv1_{br}.bias"]))
        h = F.relu(F.linear(h, sd[f"conv2_{br}.weight"], sd[f"conv2_{br}.bias"]))
        h = F.linear(h, sd[f"conv3_{br}.weight"], sd[f"conv3_{br}.bias"]).view(bs, num_obj, width)
        outs.append(torch.index_select(h[0], 0, obj[0]))
    return outs[0], outs[1]


# ----------------------------------------------------------------------------------------------
# Losses
# ----------------------------------------------------------------------------------------------
def _rotation_from_unit_quat(q: torch.Tensor) -> torch.Tensor:
    """q (P,4) unit (w,x,y,z) -> (P,3,3), term by term as lib/loss.py:18-26 (row-major 'base')."""
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rows = [
        1.0 - 2.0 * (y ** 2 + z ** 2), 2.0 * x * y - 2.0 * w * z, 2.0 * w * y + 2.0 * x * z,
        2.0 * x * y + 2.0 * z * w, 1.0 - 2.0 * (x ** 2 + z ** 2), -2.0 * w * x + 2.0 * y * z,
        -2.0 * w * y + 2.0 * x * z, 2.0 * w * x + 2.0 * y * z, 1.0 - 2.0 * (x ** 2 + y ** 2),
    ]
    return torch.stack(rows, dim=1).view(-1, 3, 3)


def _nearest_target(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """pred (P,M,3), target (M,3) -> (P,M,3): each predicted point's nearest target point, with the
    upstream call contract knn(target (1,3,M), pred (1,3,P*M)) -> 1-based (lib/loss.py:42-47)."""
    P, M = pred.shape[0], pred.shape[1]
    tgt_dm = target.t().contiguous()                          # (3,M)
    pred_dm = pred.detach().permute(2, 0, 1).contiguous().view(3, -1)
    inds = knn(tgt_dm.unsqueeze(0), pred_dm.unsqueeze(0), 1).view(-1) - 1
    sel = torch.index_select(tgt_dm, 1, inds)                 # (3,P*M)
    return sel.view(3, P, M).permute(1, 2, 0).contiguous()


def loss(pred_r, pred_t, pred_c, target, model_points, idx, points, w, refine, num_point_mesh, sym_list):
    """lib/loss.py:13-70 for bs=1.  Returns (loss, dis_at_argmax, new_points, new_target)."""
    bs, num_p, _ = pred_c.shape
    assert bs == 1
    q = pred_r / torch.norm(pred_r, dim=2).view(bs, num_p, 1)
    rot = _rotation_from_unit_quat(q.view(num_p, 4))          # ori_base
    mp = model_points.view(num_point_mesh, 3)
    tg = target.view(num_point_mesh, 3)
    t = pred_t.contiguous().view(num_p, 1, 3)
    pts = points.contiguous().view(num_p, 1, 3)
    conf = pred_c.contiguous().view(num_p)
    pred = torch.matmul(mp.unsqueeze(0), rot.transpose(2, 1)) + (pts + t)     # (P,M,3)
    if (not refine) and int(idx.view(-1)[0].item()) in sym_list:
        tgt = _nearest_target(pred, tg)
    else:
        tgt = tg.unsqueeze(0).expand(num_p, -1, -1)
    dis = torch.mean(torch.norm(pred - tgt, dim=2), dim=1)
    total = torch.mean(dis * conf - w * torch.log(conf), dim=0)
    which = torch.max(conf.view(bs, num_p), 1)[1]
    k = which[0]
    tk = t[k] + pts[k]                                         # (1,3)
    rk = rot[k].view(1, 3, 3)
    new_points = torch.bmm(pts.view(1, num_p, 3) - tk.view(1, 1, 3), rk).contiguous()
    new_target = torch.bmm(tg.view(1, num_point_mesh, 3) - tk.view(1, 1, 3), rk).contiguous()
    return total, dis[k], new_points.detach(), new_target.detach()


def loss_refine(pred_r, pred_t, target, model_points, idx, points, num_point_mesh, sym_list):
    """lib/loss_refiner.py:12-62.  Returns (dis (1,), new_points, new_target)."""
    q = pred_r.view(1, 1, -1)
    t = pred_t.view(1, 1, -1)
    n_in = points.shape[1]
    q = q / torch.norm(q, dim=2).view(1, 1, 1)
    rot = _rotation_from_unit_quat(q.view(1, 4))
    mp = model_points.view(num_point_mesh, 3)
    tg = target.view(num_point_mesh, 3)
    pred = torch.matmul(mp.unsqueeze(0), rot.transpose(2, 1)) + t.view(1, 1, 3)
    if int(idx.view(-1)[0].item()) in sym_list:
        tgt = _nearest_target(pred, tg)
    else:
        tgt = tg.unsqueeze(0)
    dis = torch.mean(torch.norm(pred - tgt, dim=2), dim=1)
    t0 = t.view(1, 1, 3)
    new_points = torch.bmm(points.view(1, n_in, 3) - t0, rot[0:1]).contiguous()
    new_target = torch.bmm(tg.view(1, num_point_mesh, 3) - t0, rot[0:1]).contiguous()
    return dis, new_points.detach(), new_target.detach()


# ----------------------------------------------------------------------------------------------
# Host pose algebra (float64) and the eval-time refine loop
# ----------------------------------------------------------------------------------------------
_EPS = np.finfo(float).eps * 4.0


def quaternion_matrix(quaternion) -> np.ndarray:
    q = np.array(quaternion, dtype=np.float64, copy=True)
    n = float(np.dot(q, q))
    if n < _EPS:
        return np.identity(4)
    q *= math.sqrt(2.0 / n)
    o = np.outer(q, q)
    m = np.identity(4)
    m[0, 0] = 1.0 - o[2, 2] - o[3, 3]; m[0, 1] = o[1, 2] - o[3, 0]; m[0, 2] = o[1, 3] + o[2, 0]
    m[1, 0] = o[1, 2] + o[3, 0]; m[1, 1] = 1.0 - o[1, 1] - o[3, 3]; m[1, 2] = o[2, 3] - o[1, 0]
    m[2, 0] = o[1, 3] - o[2, 0]; m[2, 1] = o[2, 3] + o[1, 0]; m[2, 2] = 1.0 - o[1, 1] - o[2, 2]
    return m


def quaternion_from_matrix(matrix, isprecise: bool = True) -> np.ndarray:
    """Only the isprecise=True branch is on the path (tools/eval_ycb.py:225)."""
    if not isprecise:
        raise NotImplementedError("the path only uses isprecise=True")
    M = np.asarray(matrix, dtype=np.float64)[:4, :4]
    q = np.empty(4)
    t = np.trace(M)
    if t > M[3, 3]:
        q[0] = t
        q[3] = M[1, 0] - M[0, 1]
        q[2] = M[0, 2] - M[2, 0]
        q[1] = M[2, 1] - M[1, 2]
    else:
        i, j, k = 0, 1, 2
        if M[1, 1] > M[0, 0]:
            i, j, k = 1, 2, 0
        if M[2, 2] > M[i, i]:
            i, j, k = 2, 0, 1
        t = M[i, i] - (M[j, j] + M[k, k]) + M[3, 3]
        q[i] = t
        q[j] = M[i, j] + M[j, i]
        q[k] = M[k, i] + M[i, k]
        q[3] = M[k, j] - M[j, k]
        q = q[[3, 0, 1, 2]]
    q *= 0.5 / math.sqrt(t * M[3, 3])
    if q[0] < 0.0:
        q = -q
    return q


def select_pose(pred_r, pred_t, pred_c, cloud):
    """tools/eval_ycb.py:193-201: normalise, argmax confidence, t = point + offset.
    Returns (my_r (4,) f32 ndarray, my_t (3,) f32 ndarray, which_max int)."""
    n = pred_r.shape[1]
    q = pred_r / torch.norm(pred_r, dim=2).view(1, n, 1)
    which = int(torch.max(pred_c.view(1, n), 1)[1][0].item())
    my_r = q[0][which].view(-1).detach().numpy()
    my_t = (cloud.view(n, 1, 3) + pred_t.view(n, 1, 3))[which].view(-1).detach().numpy()
    return my_r, my_t, which


def refine_pose_eval(sd_refiner: dict, cloud, emb, obj, num_obj, my_r, my_t, iterations: int):
    """tools/eval_ycb.py:205-229.  cloud (1,N,3). Returns (q (4,) f64, t (3,) f64)."""
    n = cloud.shape[1]
    my_r = np.asarray(my_r)
    my_t = np.asarray(my_t)
    for _ in range(iterations):
        T = torch.from_numpy(my_t.astype(np.float32)).view(1, 1, 3)
        my_mat = quaternion_matrix(my_r)
        R = torch.from_numpy(my_mat[:3, :3].astype(np.float32)).view(1, 3, 3)
        my_mat[0:3, 3] = my_t
        new_cloud = torch.bmm(cloud - T.expand(1, n, 3), R).contiguous()
        r2, t2 = refiner_forward(sd_refiner, new_cloud, emb, obj, num_obj)
        r2 = r2.view(1, 1, -1)
        r2 = r2 / torch.norm(r2, dim=2).view(1, 1, 1)
        m2 = quaternion_matrix(r2.view(-1).detach().numpy())
        m2[0:3, 3] = t2.view(-1).detach().numpy()
        final = np.dot(my_mat, m2)
        rot_only = final.copy()
        rot_only[0:3, 3] = 0
        my_r = quaternion_from_matrix(rot_only, True)
        my_t = np.array([final[0][3], final[1][3], final[2][3]])
    return np.asarray(my_r, dtype=np.float64), np.asarray(my_t, dtype=np.float64)


def estimate_and_refine(sd_est, sd_ref, img, cloud, choose, obj, num_obj, iterations=2):
    """One pose, eval_ycb semantics: returns 7 floats [qw qx qy qz tx ty tz] (f64)."""
    with torch.no_grad():
        r, t, c, emb = posenet_forward(sd_est, img, cloud, choose, obj, num_obj)
        my_r, my_t, _ = select_pose(r, t, c, cloud)
        q, tt = refine_pose_eval(sd_ref, cloud, emb, obj, num_obj, my_r, my_t, iterations)
    return np.concatenate([q, tt])


# ----------------------------------------------------------------------------------------------
# Training step (tools/train.py:143-169): per-sample forward + backward, gradients accumulate (SUM)
# ----------------------------------------------------------------------------------------------
def posenet_state_shapes(num_obj: int) -> dict:
    """Parameter names and shapes of the reference PoseNet's state_dict (lib/network.py:27-37,40-49,73-91; lib/pspnet.py:40-60;
    lib/extractors.py:78-103), as a static table: lets a caller build synthetic weights without constructing any module."""
    sh = {"cnn.model.module.feats.conv1.weight": (64, 3, 7, 7)}
    cin = 64
    for li, cout in enumerate((64, 128, 256, 512), 1):
        for bi in (0, 1):
            pre = f"cnn.model.module.feats.layer{li}.{bi}."
            sh[pre + "conv1.weight"] = (cout, cin if bi == 0 else cout, 3, 3)
            sh[pre + "conv2.weight"] = (cout, cout, 3, 3)
            if bi == 0 and cin != cout:
                sh[pre + "downsample.0.weight"] = (cout, cin, 1, 1)
        cin = cout
    for i in range(4):
        sh[f"cnn.model.module.psp.stages.{i}.1.weight"] = (512, 512, 1, 1)
    sh["cnn.model.module.psp.bottleneck.weight"] = (1024, 2560, 1, 1)
    sh["cnn.model.module.psp.bottleneck.bias"] = (1024,)
    for name, co, ci in (("up_1", 256, 1024), ("up_2", 64, 256), ("up_3", 64, 64)):
        sh[f"cnn.model.module.{name}.conv.1.weight"] = (co, ci, 3, 3)
        sh[f"cnn.model.module.{name}.conv.1.bias"] = (co,)
        sh[f"cnn.model.module.{name}.conv.2.weight"] = (1,)
    sh["cnn.model.module.final.0.weight"] = (32, 64, 1, 1)
    sh["cnn.model.module.final.0.bias"] = (32,)
    sh["cnn.model.module.classifier.0.weight"] = (256, 256)
    sh["cnn.model.module.classifier.0.bias"] = (256,)
    sh["cnn.model.module.classifier.2.weight"] = (21, 256)
    sh["cnn.model.module.classifier.2.bias"] = (21,)
    for name, co, ci in (("conv1", 64, 3), ("conv2", 128, 64), ("e_conv1", 64, 32), ("e_conv2", 128, 64), ("conv5", 512, 256),
                         ("conv6", 1024, 512)):
        sh[f"feat.{name}.weight"] = (co, ci, 1)
        sh[f"feat.{name}.bias"] = (co,)
    for layer, co, ci in ((1, 640, 1408), (2, 256, 640), (3, 128, 256)):
        for b in "rtc":
            sh[f"conv{layer}_{b}.weight"] = (co, ci, 1)
            sh[f"conv{layer}_{b}.bias"] = (co,)
    for b, k in (("r", 4), ("t", 3), ("c", 1)):
        sh[f"conv4_{b}.weight"] = (num_obj * k, 128, 1)
        sh[f"conv4_{b}.bias"] = (num_obj * k,)
    return sh


def refiner_state_shapes(num_obj: int) -> dict:
    """The same for PoseRefineNet (lib/network.py:134-146,170-183)."""
    sh = {}
    for name, co, ci in (("conv1", 64, 3), ("conv2", 128, 64), ("e_conv1", 64, 32), ("e_conv2", 128, 64), ("conv5", 512, 384),
                         ("conv6", 1024, 512)):
        sh[f"feat.{name}.weight"] = (co, ci, 1)
        sh[f"feat.{name}.bias"] = (co,)
    for b, k in (("r", 4), ("t", 3)):
        sh[f"conv1_{b}.weight"], sh[f"conv1_{b}.bias"] = (512, 1024), (512,)
        sh[f"conv2_{b}.weight"], sh[f"conv2_{b}.bias"] = (128, 512), (128,)
        sh[f"conv3_{b}.weight"], sh[f"conv3_{b}.bias"] = (num_obj * k, 128), (num_obj * k,)
    return sh


def _leaf_state_dict(sd: dict) -> dict:
    return {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}


def estimator_gradients(sd_est: dict, crops, num_obj, num_point_mesh, sym_list, w):
    """Estimator phase (tools/train.py:152-153, :161): for every crop (bs=1 dict with img, points, choose,
    target, model_points, idx) loss.backward(); returns ({name: summed grad or None}, [loss], [dis])."""
    leaf = _leaf_state_dict(sd_est)
    losses, dists = [], []
    for d in crops:
        r, t, c, _ = posenet_forward(leaf, d["img"], d["points"], d["choose"], d["idx"], num_obj)
        total, dis, _, _ = loss(r, t, c, d["target"], d["model_points"], d["idx"], d["points"], w, False,
                                num_point_mesh, sym_list)
        total.backward()
        losses.append(float(total))
        dists.append(float(dis))
    return {k: v.grad for k, v in leaf.items()}, losses, dists


def refiner_gradients(sd_est: dict, sd_ref: dict, crops, num_obj, num_point_mesh, sym_list, w, iteration):
    """Refiner phase (tools/train.py:152-159): Loss(..., refine=True) then `iteration` x (refiner, Loss_refine,
    dis.backward()); only the refiner's gradients matter (tools/train.py:93)."""
    leaf = _leaf_state_dict(sd_ref)
    dists = []
    for d in crops:
        with torch.no_grad():
            r, t, c, emb = posenet_forward(sd_est, d["img"], d["points"], d["choose"], d["idx"], num_obj)
            _, _, pts, tgt = loss(r, t, c, d["target"], d["model_points"], d["idx"], d["points"], w, True,
                                  num_point_mesh, sym_list)
        for _ in range(iteration):
            rr, tt = refiner_forward(leaf, pts, emb, d["idx"], num_obj)
            dis, pts, tgt = loss_refine(rr, tt, tgt, d["model_points"], d["idx"], pts, num_point_mesh, sym_list)
            dis.backward()
        dists.append(float(dis))
    return {k: v.grad for k, v in leaf.items()}, dists


def adam_reference(params: dict, grads: dict, state: dict, lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam (the reference's optimiser, tools/train.py:97) applied functionally: params / state are
    updated in place; parameters whose gradient is None are skipped, as torch does."""
    state["step"] = state.get("step", 0) + 1
    k = state["step"]
    for name, p in params.items():
        g = grads.get(name)
        if g is None:
            continue
        m = state.setdefault("m." + name, torch.zeros_like(p))
        v = state.setdefault("v." + name, torch.zeros_like(p))
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        denom = (v.sqrt() / math.sqrt(1 - betas[1] ** k)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / (1 - betas[0] ** k))
    return params


# ----------------------------------------------------------------------------------------------
# Input preparation (tools/eval_ycb.py:54-91, :147-190), numpy like the reference
# ----------------------------------------------------------------------------------------------
_BORDER_LIST = [-1, 40, 80, 120, 160, 200, 240, 280, 320, 360, 400, 440, 480, 520, 560, 600, 640, 680]


def crop_bbox(roi, img_width=480, img_length=640):
    """tools/eval_ycb.py:54-91 for one PoseCNN row."""
    rmin, rmax = int(roi[3]) + 1, int(roi[5]) - 1
    cmin, cmax = int(roi[2]) + 1, int(roi[4]) - 1
    ext = []
    for e in (rmax - rmin, cmax - cmin):
        for tt in range(len(_BORDER_LIST) - 1):
            if _BORDER_LIST[tt] < e < _BORDER_LIST[tt + 1]:
                e = _BORDER_LIST[tt + 1]
                break
        ext.append(e)
    center = [int((rmin + rmax) / 2), int((cmin + cmax) / 2)]
    rmin, rmax = center[0] - int(ext[0] / 2), center[0] + int(ext[0] / 2)
    cmin, cmax = center[1] - int(ext[1] / 2), center[1] + int(ext[1] / 2)
    if rmin < 0:
        rmax += -rmin
        rmin = 0
    if cmin < 0:
        cmax += -cmin
        cmin = 0
    if rmax > img_width:
        rmin -= rmax - img_width
        rmax = img_width
    if cmax > img_length:
        cmin -= cmax - img_length
        cmax = img_length
    return rmin, rmax, cmin, cmax


def build_crop(img_u8, depth, label, roi, itemid, num_points, cam=(312.9869, 241.3109, 1066.778, 1067.487, 10000.0),
               rng=None):
    """tools/eval_ycb.py:150-178 for one object.  img_u8 (H,W,3) uint8, depth / label (H,W).  Returns
    (cloud (N,3) f32, choose (N,) int64, img_masked (3,h,w) f32 normalised, count).  With more than num_points masked
    pixels the reference shuffles with the global numpy RNG; pass `rng` (np.random.RandomState) to reproduce that."""
    import numpy.ma as ma
    H, W = depth.shape
    xmap = np.array([[j for _ in range(W)] for j in range(H)])
    ymap = np.array([[i for i in range(W)] for _ in range(H)])
    cam_cx, cam_cy, cam_fx, cam_fy, cam_scale = cam
    rmin, rmax, cmin, cmax = crop_bbox(roi, H, W)
    mask = ma.getmaskarray(ma.masked_equal(label, itemid)) * ma.getmaskarray(ma.masked_not_equal(depth, 0))
    choose = mask[rmin:rmax, cmin:cmax].flatten().nonzero()[0]
    count = len(choose)
    if count > num_points:
        c_mask = np.zeros(count, dtype=int)
        c_mask[:num_points] = 1
        (rng or np.random).shuffle(c_mask)
        choose = choose[c_mask.nonzero()]
    else:
        choose = np.pad(choose, (0, num_points - count), 'wrap')
    depth_masked = depth[rmin:rmax, cmin:cmax].flatten()[choose][:, np.newaxis].astype(np.float32)
    xmap_masked = xmap[rmin:rmax, cmin:cmax].flatten()[choose][:, np.newaxis].astype(np.float32)
    ymap_masked = ymap[rmin:rmax, cmin:cmax].flatten()[choose][:, np.newaxis].astype(np.float32)
    pt2 = depth_masked / np.float32(cam_scale)
    pt0 = (ymap_masked - np.float32(cam_cx)) * pt2 / np.float32(cam_fx)
    pt1 = (xmap_masked - np.float32(cam_cy)) * pt2 / np.float32(cam_fy)
    cloud = np.concatenate((pt0, pt1, pt2), axis=1).astype(np.float32)
    img_masked = np.transpose(np.array(img_u8)[:, :, :3], (2, 0, 1))[:, rmin:rmax, cmin:cmax].astype(np.float32)
    mean = np.array([0.485, 0.456, 0.406], dtype=np.float32).reshape(3, 1, 1)
    std = np.array([0.229, 0.224, 0.225], dtype=np.float32).reshape(3, 1, 1)
    img_masked = (img_masked - mean) / std                      # transforms.Normalize: sub_(mean).div_(std) in fp32
    return cloud, choose.astype(np.int64), img_masked, count
