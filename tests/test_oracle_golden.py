"""The oracle (oracle/df_oracle.py, oracle/knn_ref.c) pinned against
  * the reference's own known-answer doctests for the two quaternion helpers
    (lib/transformations.py:1257-1265, :1287-1317), and
  * outputs of the reference itself on seeded inputs (tests/golden/*.npz, made by
    tests/golden/make_golden.py in the build container).
CPU only."""
import math

import numpy as np
import pytest
import torch

from conftest import golden
from densefusion_b200 import synth
from oracle import df_oracle as O
from util import rel


# ---- quaternion KATs (reference doctests) -----------------------------------------------------
def _rotation_matrix(angle, direction):
    d = np.asarray(direction, dtype=np.float64)
    d = d / np.linalg.norm(d)
    s, c = math.sin(angle), math.cos(angle)
    R = np.diag([c, c, c]) + np.outer(d, d) * (1.0 - c)
    d = d * s
    R += np.array([[0.0, -d[2], d[1]], [d[2], 0.0, -d[0]], [-d[1], d[0], 0.0]])
    M = np.identity(4)
    M[:3, :3] = R
    return M


def test_quaternion_matrix_kats():
    assert np.allclose(O.quaternion_matrix([0.99810947, 0.06146124, 0, 0]), _rotation_matrix(0.123, [1, 0, 0]))
    assert np.allclose(O.quaternion_matrix([1, 0, 0, 0]), np.identity(4))
    assert np.allclose(O.quaternion_matrix([0, 1, 0, 0]), np.diag([1, -1, -1, 1]))


def test_quaternion_from_matrix_kats():
    assert np.allclose(O.quaternion_from_matrix(np.identity(4), True), [1, 0, 0, 0])
    R = _rotation_matrix(0.123, (1, 2, 3))
    assert np.allclose(O.quaternion_from_matrix(R, True), [0.9981095, 0.0164262, 0.0328524, 0.0492786])
    # round trip through every pivot branch (trace<=1 with each diagonal dominant)
    for q in ([0.1, 0.9, 0.2, 0.3], [0.1, 0.2, 0.9, 0.3], [0.1, 0.2, 0.3, 0.9], [0.9, 0.1, 0.2, 0.3]):
        q = np.array(q) / np.linalg.norm(q)
        back = O.quaternion_from_matrix(O.quaternion_matrix(q), True)
        assert np.allclose(back, q, atol=1e-12)


# ---- kNN emulator -----------------------------------------------------------------------------
def test_knn_c_vs_numpy_and_general_k():
    g = torch.Generator().manual_seed(5)
    ref = torch.randn(1, 3, 37, generator=g)
    qry = torch.randn(1, 3, 53, generator=g)
    a = O.knn(ref, qry, 1).view(-1).numpy()
    b = O.knn_numpy_small(ref[0].numpy(), qry[0].numpy())
    assert np.array_equal(a, b)
    # general-k path agrees with the k=1 fast path and returns a (distance, index)-sorted list
    k3 = O.knn(ref, qry, 3)
    assert np.array_equal(k3[0, 0].numpy(), a)
    d = ((ref[0].t()[None] - qry[0].t()[:, None]) ** 2).sum(-1)      # (Q,R)
    order = torch.argsort(d, dim=1, stable=True)[:, :3].t() + 1
    assert np.array_equal(k3[0].numpy(), order.numpy())


def test_knn_ties_lowest_index_and_nan():
    ref = torch.tensor([[[0.0, 1.0, 1.0, 0.0], [0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0]]])
    qry = torch.tensor([[[0.5, 1.0], [0.0, 0.0], [0.0, 0.0]]])
    out = O.knn(ref, qry, 1).view(-1).tolist()
    assert out == [1, 2]                       # 0.5 ties refs 1,2,3,4 pairwise -> first wins; exact hit -> 2
    ref[0, 0, 0] = float("nan")               # NaN in row 0: nothing is ever '<' NaN -> stays 1
    assert O.knn(ref, qry, 1).view(-1).tolist() == [1, 1]


def test_ply_pair_add_adds():
    g = golden("ply_pair")
    pred, tgt = g["pred"], g["target"]
    assert abs(np.mean(np.linalg.norm(pred - tgt, axis=1)) - 0.0168566) < 1e-7
    p32, t32 = torch.from_numpy(pred.astype(np.float32)), torch.from_numpy(tgt.astype(np.float32))
    inds = O.knn(t32.t().contiguous()[None], p32.t().contiguous()[None], 1).view(-1) - 1
    assert np.array_equal(inds.numpy().astype(np.int16), g["inds"])
    adds = torch.mean(torch.norm(p32 - t32[inds], dim=1)).item()
    assert abs(adds - 0.0092865) < 1e-7


# ---- network / loss / refine against the reference's own outputs ---------------------------
def _nets(g):
    from densefusion_b200.lib.network import PoseNet, PoseRefineNet   # only for key/shape listing
    case, n, o, m, h, w, obj, seed, iters = [int(v) for v in g["meta"]]
    est_sd = synth.synth_state_dict(synth.shapes_of(PoseNet(n, o)), seed)
    ref_sd = synth.synth_state_dict(synth.shapes_of(PoseRefineNet(n, o)), seed + 1)
    d = synth.synth_crop(case, n, m, o, (h, w), obj)
    return est_sd, ref_sd, d, (n, o, m, iters)


@pytest.mark.parametrize("name", ["c0_linemod_add", "c1_ycb_adds"])
def test_full_path_vs_reference(name):
    g = golden(name)
    est_sd, ref_sd, d, (n, o, m, iters) = _nets(g)
    sym = [int(v) for v in g["sym_list"]]
    w = float(g["w"])
    with torch.no_grad():
        r, t, c, emb = O.posenet_forward(est_sd, d["img"], d["points"], d["choose"], d["idx"], o)
    assert rel(emb, g["emb"]) < 1e-5
    assert rel(r, g["pred_r"]) < 1e-4 and rel(t, g["pred_t"]) < 1e-4 and rel(c, g["pred_c"]) < 1e-5
    assert int(torch.max(c.view(1, -1), 1)[1][0]) == int(g["which_max"])
    # loss on the reference's predictions (isolates the loss restatement), fwd + grads
    pr, pt, pc = [torch.from_numpy(g[k]).clone().requires_grad_(True) for k in ("pred_r", "pred_t", "pred_c")]
    loss, dis, npts, ntgt = O.loss(pr, pt, pc, d["target"], d["model_points"], d["idx"], d["points"], w, False, m, sym)
    loss.backward()
    assert rel(loss.item(), g["loss"]) < 1e-5 and rel(dis.item(), g["dis"]) < 1e-5
    assert rel(npts, g["new_points"]) < 1e-5 and rel(ntgt, g["new_target"]) < 1e-5
    assert rel(pr.grad, g["g_pred_r"]) < 1e-4 and rel(pt.grad, g["g_pred_t"]) < 1e-4 and rel(pc.grad, g["g_pred_c"]) < 1e-4
    with torch.no_grad():
        l2, d2, _, _ = O.loss(pr, pt, pc, d["target"], d["model_points"], d["idx"], d["points"], w, True, m, sym)
    assert rel(l2.item(), g["loss_refineflag"]) < 1e-5 and rel(d2.item(), g["dis_refineflag"]) < 1e-5
    # training-style refine chain
    pts, tgt = npts, ntgt
    emb_ref = torch.from_numpy(g["emb"])
    for it in range(iters):
        with torch.no_grad():
            rr, tt = O.refiner_forward(ref_sd, pts, emb_ref, d["idx"], o)
        assert rel(rr, g[f"train_r{it}"]) < 1e-4 and rel(tt, g[f"train_t{it}"]) < 1e-4
        rr_l, tt_l = rr.clone().requires_grad_(True), tt.clone().requires_grad_(True)
        dis_r, pts, tgt = O.loss_refine(rr_l, tt_l, tgt, d["model_points"], d["idx"], pts, m, sym)
        dis_r.backward()
        assert rel(dis_r, g[f"train_dis{it}"]) < 1e-4
        assert rel(pts, g[f"train_pts{it}"]) < 1e-4 and rel(tgt, g[f"train_tgt{it}"]) < 1e-4
        assert rel(rr_l.grad, g[f"train_g_r{it}"]) < 1e-3 and rel(tt_l.grad, g[f"train_g_t{it}"]) < 1e-3
    # eval-style pose
    with torch.no_grad():
        my_r, my_t, which = O.select_pose(torch.from_numpy(g["pred_r"]), torch.from_numpy(g["pred_t"]),
                                          torch.from_numpy(g["pred_c"]), d["points"])
        assert which == int(g["which_max"])
        assert rel(np.append(my_r, my_t), g["pose0"]) < 1e-6
        for n_it in (1, iters, 4):
            q, tt = O.refine_pose_eval(ref_sd, d["points"], emb_ref, d["idx"], o, my_r, my_t, n_it)
            assert rel(np.append(q, tt), g[f"pose_iter{n_it}"]) < 1e-4


@pytest.mark.parametrize("name", ["loss_add_n500_m500", "loss_adds_n500_m500", "loss_adds_n100_m2600",
                                  "loss_add_n1000_m500"])
def test_loss_only_vs_reference(name):
    g = golden(name)
    case, n, m, obj = [int(v) for v in g["meta"]]
    sym = [int(v) for v in g["sym_list"]]
    w = float(g["w"])
    d = synth.synth_crop(case, n, m, 21, (40, 40), obj)
    pred_r, pred_t, _ = synth.synth_predictions(case, n)
    pr, pt = pred_r.clone().requires_grad_(True), pred_t.clone().requires_grad_(True)
    pc = torch.from_numpy(g["pred_c_tied"]).clone().requires_grad_(True)
    loss, dis, npts, ntgt = O.loss(pr, pt, pc, d["target"], d["model_points"], d["idx"], d["points"], w, False, m, sym)
    loss.backward()
    assert rel(loss.item(), g["loss"]) < 1e-5 and rel(dis.item(), g["dis"]) < 1e-5
    assert rel(npts, g["new_points"]) < 1e-5 and rel(ntgt, g["new_target"]) < 1e-5
    assert rel(pr.grad, g["g_pred_r"]) < 1e-4 and rel(pt.grad, g["g_pred_t"]) < 1e-4 and rel(pc.grad, g["g_pred_c"]) < 1e-4
    r1 = pred_r[0, 5].view(1, 4).clone().requires_grad_(True)
    t1 = (pred_t[0, 5] + torch.tensor([0.0, 0.0, 0.8])).view(1, 3).clone().requires_grad_(True)
    dis_r, np_r, nt_r = O.loss_refine(r1, t1, d["target"], d["model_points"], d["idx"], d["points"], m, sym)
    dis_r.backward()
    assert rel(dis_r, g["ref_dis"]) < 1e-5
    assert rel(np_r, g["ref_new_points"]) < 1e-5 and rel(nt_r, g["ref_new_target"]) < 1e-5
    assert rel(r1.grad, g["ref_g_r"]) < 1e-4 and rel(t1.grad, g["ref_g_t"]) < 1e-4


def test_oracle_on_real_geometry_customcad_triple():
    """The oracle against the reference's own Loss / Loss_refine / transformations on the real-geometry triple of the reference
    tree (datasets/customCAD/*.ply, golden generated by tests/golden/make_golden.py::customcad_case)."""
    g = golden("customcad_triple")
    cloud, model, target = [torch.from_numpy(g[k]) for k in ("cloud", "model", "target")]
    pr, pt, pc = [torch.from_numpy(g[k]) for k in ("pred_r", "pred_t", "pred_c")]
    m = model.shape[1]
    sym = [int(v) for v in g["sym_list"]]
    for tag, obj in (("add", 3), ("adds", 5)):
        idx = torch.tensor([[obj]])
        loss, dis, npts, ntgt = O.loss(pr, pt, pc, target, model, idx, cloud, float(g["w"]), False, m, sym)
        assert rel(loss, g[f"{tag}_loss"]) < 1e-6 and rel(dis, g[f"{tag}_dis"]) < 1e-6
        assert rel(npts, g[f"{tag}_new_points"]) < 1e-6 and rel(ntgt, g[f"{tag}_new_target"]) < 1e-6
        r1 = (torch.tensor([1.0, 0.01, -0.02, 0.015]) * 1.3).view(1, 4)
        t1 = torch.tensor([[0.002, -0.001, 0.003]])
        dis_r, np_r, nt_r = O.loss_refine(r1, t1, ntgt, model, idx, npts, m, sym)
        assert rel(dis_r, g[f"{tag}_ref_dis"]) < 1e-6 and rel(np_r, g[f"{tag}_ref_new_points"]) < 1e-6
    my_r, my_t, which = O.select_pose(pr, pt, pc, cloud)
    assert which == int(g["which"]) and np.allclose(np.append(my_r, my_t), g["pose0"], atol=1e-7)
