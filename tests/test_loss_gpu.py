"""K3 parity: fused loss / selection / backward against the reference's outputs (tests/golden) and the oracle."""
import numpy as np
import pytest
import torch

from conftest import golden
from densefusion_b200 import synth
from oracle import df_oracle as O
from util import rel

pytestmark = pytest.mark.gpu
TOL = 1e-4        # BASELINE.json: poses and losses within 1e-4 relative in fp32


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


@pytest.mark.parametrize("name", ["loss_add_n500_m500", "loss_adds_n500_m500", "loss_adds_n100_m2600",
                                  "loss_add_n1000_m500"])
def test_loss_and_loss_refine_vs_reference_golden(name):
    from densefusion_b200.lib.loss import Loss
    from densefusion_b200.lib.loss_refiner import Loss_refine
    g = golden(name)
    case, n, m, obj = [int(v) for v in g["meta"]]
    sym = [int(v) for v in g["sym_list"]]
    w = float(g["w"])
    d = _cuda(synth.synth_crop(case, n, m, 21, (40, 40), obj))
    pred_r, pred_t, _ = synth.synth_predictions(case, n)
    pr = pred_r.cuda().requires_grad_(True)
    pt = pred_t.cuda().requires_grad_(True)
    pc = torch.from_numpy(g["pred_c_tied"]).cuda().requires_grad_(True)
    loss, dis, npts, ntgt = Loss(m, sym)(pr, pt, pc, d["target"], d["model_points"], d["idx"], d["points"], w, False)
    assert loss.dim() == 0 and dis.dim() == 0 and tuple(npts.shape) == (1, n, 3) and tuple(ntgt.shape) == (1, m, 3)
    assert not npts.requires_grad and not ntgt.requires_grad
    loss.backward()
    assert rel(loss, g["loss"]) < TOL and rel(dis, g["dis"]) < TOL
    assert rel(npts, g["new_points"]) < TOL and rel(ntgt, g["new_target"]) < TOL
    assert rel(pr.grad, g["g_pred_r"]) < 1e-3 and rel(pt.grad, g["g_pred_t"]) < 1e-3 and rel(pc.grad, g["g_pred_c"]) < 1e-3
    r1 = pred_r[0, 5].view(1, 4).cuda().requires_grad_(True)
    t1 = (pred_t[0, 5] + torch.tensor([0.0, 0.0, 0.8])).view(1, 3).cuda().requires_grad_(True)
    dis_r, np_r, nt_r = Loss_refine(m, sym)(r1, t1, d["target"], d["model_points"], d["idx"], d["points"])
    assert tuple(dis_r.shape) == (1,)
    dis_r.backward()
    assert rel(dis_r, g["ref_dis"]) < TOL
    assert rel(np_r, g["ref_new_points"]) < TOL and rel(nt_r, g["ref_new_target"]) < TOL
    assert rel(r1.grad, g["ref_g_r"]) < 1e-3 and rel(t1.grad, g["ref_g_t"]) < 1e-3


@pytest.mark.parametrize("name", ["c0_linemod_add", "c1_ycb_adds"])
def test_loss_on_reference_predictions(name):
    """The reference's own head outputs in, the reference's loss / selection / gradients out."""
    from densefusion_b200.lib.loss import Loss
    g = golden(name)
    case, n, o, m, h, w_, obj, seed, iters = [int(v) for v in g["meta"]]
    sym = [int(v) for v in g["sym_list"]]
    w = float(g["w"])
    d = _cuda(synth.synth_crop(case, n, m, o, (h, w_), obj))
    pr, pt, pc = [torch.from_numpy(g[k]).cuda().requires_grad_(True) for k in ("pred_r", "pred_t", "pred_c")]
    crit = Loss(m, sym)
    loss, dis, npts, ntgt = crit(pr, pt, pc, d["target"], d["model_points"], d["idx"], d["points"], w, False)
    loss.backward()
    assert rel(loss, g["loss"]) < TOL and rel(dis, g["dis"]) < TOL
    assert rel(npts, g["new_points"]) < TOL and rel(ntgt, g["new_target"]) < TOL
    assert rel(pr.grad, g["g_pred_r"]) < 1e-3 and rel(pt.grad, g["g_pred_t"]) < 1e-3 and rel(pc.grad, g["g_pred_c"]) < 1e-3
    with torch.no_grad():
        l2, d2, _, _ = crit(pr, pt, pc, d["target"], d["model_points"], d["idx"], d["points"], w, True)
    assert rel(l2, g["loss_refineflag"]) < TOL and rel(d2, g["dis_refineflag"]) < TOL


def test_selection_and_fused_nn_are_bit_exact():
    """argmax index == torch.max first index; the fused kernel's nearest-target indices == the oracle kNN run on
    the kernel's own transformed points (identical inputs -> identical indices)."""
    from densefusion_b200 import ops
    n, m = 500, 500
    d = synth.synth_crop(21, n, m, 21, (40, 40), 15)
    pred_r, pred_t, pred_c = synth.synth_predictions(21, n)
    pred_c[0, 123, 0] = 0.99
    pred_c[0, 400, 0] = 0.99
    dc = _cuda(d)
    st = ops.loss_forward(pred_r.cuda(), pred_t.cuda(), pred_c.cuda(), dc["target"], dc["model_points"], dc["points"],
                          dc["points"], dc["idx"], ops.sym_mask(synth.YCB_SYM), True, 0.015, debug=True)
    assert int(st.which[0]) == int(torch.max(pred_c.view(1, -1), 1)[1][0]) == 123
    pred = st.dbg_pred[0].cpu()                                  # (P,M,3)
    want = O.knn(d["target"][0].t().contiguous()[None], pred.permute(2, 0, 1).reshape(3, -1)[None], 1).view(n, m) - 1
    assert torch.equal(st.dbg_nn[0].cpu().long(), want)
    # and the transformed points themselves match the reference formula
    q = pred_r / pred_r.norm(dim=2, keepdim=True)
    rot = O._rotation_from_unit_quat(q.view(n, 4))
    ref_pred = torch.matmul(d["model_points"][0][None], rot.transpose(2, 1)) + (d["points"].view(n, 1, 3) + pred_t.view(n, 1, 3))
    assert rel(pred, ref_pred) < 1e-6


def test_loss_batched_equals_per_crop_and_oracle():
    from densefusion_b200.lib.loss import Loss
    from densefusion_b200.lib.loss_refiner import Loss_refine
    n, m, B = 500, 500, 5
    objs = [3, 12, 15, 0, 20]
    crops = [synth.synth_crop(30 + i, n, m, 21, (40, 40), objs[i]) for i in range(B)]
    preds = [synth.synth_predictions(30 + i, n) for i in range(B)]
    cat = lambda k: torch.cat([c[k] for c in crops], 0).cuda()
    pr = torch.cat([p[0] for p in preds], 0).cuda()
    pt = torch.cat([p[1] for p in preds], 0).cuda()
    pc = torch.cat([p[2] for p in preds], 0).cuda()
    crit = Loss(m, synth.YCB_SYM)
    loss, dis, npts, ntgt = crit(pr, pt, pc, cat("target"), cat("model_points"), cat("idx"), cat("points"), 0.015, False)
    assert tuple(loss.shape) == (B,)
    for i in range(B):
        c = crops[i]
        l1, d1, np1, nt1 = O.loss(preds[i][0], preds[i][1], preds[i][2], c["target"], c["model_points"], c["idx"],
                                  c["points"], 0.015, False, m, synth.YCB_SYM)
        assert rel(loss[i], l1) < TOL and rel(dis[i], d1) < TOL
        assert rel(npts[i], np1[0]) < TOL and rel(ntgt[i], nt1[0]) < TOL
        li, di, _, _ = crit(pr[i:i + 1], pt[i:i + 1], pc[i:i + 1], c["target"].cuda(), c["model_points"].cuda(),
                            c["idx"].cuda(), c["points"].cuda(), 0.015, False)
        assert float(li) == float(loss[i]) and float(di) == float(dis[i])       # deterministic, batch-invariant
    # batched Loss_refine
    r1 = pr[:, 7].contiguous()
    t1 = (pt[:, 7] + torch.tensor([0.0, 0.0, 0.8], device="cuda")).contiguous()
    dis_r, _, _ = Loss_refine(m, synth.YCB_SYM)(r1, t1, cat("target"), cat("model_points"), cat("idx"), cat("points"))
    for i in range(B):
        c = crops[i]
        want, _, _ = O.loss_refine(r1[i:i + 1].cpu(), t1[i:i + 1].cpu(), c["target"], c["model_points"], c["idx"],
                                   c["points"], m, synth.YCB_SYM)
        assert rel(dis_r[i], want) < TOL


def test_loss_is_run_to_run_deterministic():
    from densefusion_b200 import ops
    n, m = 500, 500
    d = _cuda(synth.synth_crop(41, n, m, 21, (40, 40), 12))
    pr, pt, pc = [t.cuda() for t in synth.synth_predictions(41, n)]
    outs = []
    for _ in range(3):
        st = ops.loss_forward(pr, pt, pc, d["target"], d["model_points"], d["points"], d["points"], d["idx"],
                              ops.sym_mask(synth.YCB_SYM), True, 0.015)
        outs.append((float(st.loss[0]), float(st.dis_sel[0]), st.new_points.clone()))
    assert outs[0][0] == outs[1][0] == outs[2][0] and outs[0][1] == outs[1][1]
    assert torch.equal(outs[0][2], outs[2][2])


def test_real_geometry_customcad_triple_vs_reference():
    """Real geometry from the reference tree (datasets/customCAD/{depth_projected,model,target}.ply: observed cloud, CAD model,
    model under the ground-truth pose).  Golden values come from the reference's OWN Loss / Loss_refine / transformations
    (tests/golden/make_golden.py::customcad_case): K3 (ADD and, kNN-based, ADD-S) with gradients, K4 through the symmetric
    branch, and K5 (selection, re-expression of the cloud, pose composition)."""
    from densefusion_b200 import ops
    from densefusion_b200.lib.loss import Loss
    from densefusion_b200.lib.loss_refiner import Loss_refine
    g = golden("customcad_triple")
    cloud, model, target = [torch.from_numpy(g[k]).cuda() for k in ("cloud", "model", "target")]
    n, m = cloud.shape[1], model.shape[1]
    sym = [int(v) for v in g["sym_list"]]
    for tag, obj in (("add", 3), ("adds", 5)):
        idx = torch.tensor([[obj]], device="cuda")
        pr, pt, pc = [torch.from_numpy(g[k]).cuda().requires_grad_(True) for k in ("pred_r", "pred_t", "pred_c")]
        loss, dis, npts, ntgt = Loss(m, sym)(pr, pt, pc, target, model, idx, cloud, float(g["w"]), False)
        loss.backward()
        assert rel(loss, g[f"{tag}_loss"]) < TOL and rel(dis, g[f"{tag}_dis"]) < TOL
        assert rel(npts, g[f"{tag}_new_points"]) < TOL and rel(ntgt, g[f"{tag}_new_target"]) < TOL
        # ADD-S: the 500 000 nearest-target assignments are made on the kernel's own fp32 transformed points; on real geometry a
        # few near-ties resolve differently from the reference's CPU points and nothing averages them out in a per-hypothesis
        # gradient (same effect and bound as test_training_gpu.TOL_HEAD_SINGLE_ADDS); plain ADD keeps the strict bound
        gtol = 1e-3 if tag == "add" else 5e-3
        errs = (rel(pr.grad, g[f"{tag}_g_r"]), rel(pt.grad, g[f"{tag}_g_t"]), rel(pc.grad, g[f"{tag}_g_c"]))
        print(f"customCAD {tag}: gradient errors {errs}")
        assert max(errs) < gtol, (tag, errs)
        r1 = (torch.tensor([1.0, 0.01, -0.02, 0.015]) * 1.3).view(1, 4).cuda().requires_grad_(True)
        t1 = torch.tensor([[0.002, -0.001, 0.003]], device="cuda", requires_grad=True)
        dis_r, np_r, nt_r = Loss_refine(m, sym)(r1, t1, ntgt, model, idx, npts)
        dis_r.backward()
        assert rel(dis_r, g[f"{tag}_ref_dis"]) < TOL
        assert rel(np_r, g[f"{tag}_ref_new_points"]) < TOL and rel(nt_r, g[f"{tag}_ref_new_target"]) < TOL
        assert rel(r1.grad, g[f"{tag}_ref_g_r"]) < gtol and rel(t1.grad, g[f"{tag}_ref_g_t"]) < gtol
    # K5: confidence argmax + pose, re-expressed cloud, one composition (tools/eval_ycb.py:193-229 on the reference's functions)
    pr, pt, pc = [torch.from_numpy(g[k]).cuda() for k in ("pred_r", "pred_t", "pred_c")]
    pose, which = ops.select_pose(pr, pt, pc, cloud)
    assert int(which[0]) == int(g["which"])
    assert np.allclose(pose[0].cpu().numpy(), g["pose0"].astype(np.float64), rtol=0, atol=1e-6)
    assert rel(ops.cloud_transform(cloud, pose), g["new_cloud"]) < 1e-5
    ops.pose_compose_(pose, torch.from_numpy(g["r2"]).view(1, 4).cuda(), torch.from_numpy(g["t2"]).view(1, 3).cuda())
    assert np.allclose(pose[0].cpu().numpy(), g["pose1"], rtol=0, atol=2e-6), (pose[0].cpu().numpy(), g["pose1"])
