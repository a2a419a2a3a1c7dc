"""Result record + accuracy metrics.  CPU: .mat round trip and the VOCap restatement against a
direct numpy transcription of the MATLAB steps; GPU: pose distances against the reference-derived goldens."""
import numpy as np
import pytest
import torch

from conftest import golden


def test_poses_mat_roundtrip(tmp_path):
    from densefusion_b200 import evaluation as E
    poses = np.random.default_rng(0).normal(size=(5, 7))
    poses[2] = 0.0                                              # lost detection (tools/eval_ycb.py:234-237)
    p = str(tmp_path / "0000.mat")
    E.save_poses_mat(p, poses)
    import scipy.io as scio
    raw = scio.loadmat(p)
    assert "poses" in raw and raw["poses"].shape == (5, 7) and raw["poses"].dtype == np.float64
    assert np.array_equal(E.load_poses_mat(p), poses)


def _vocap_matlab(D, max_distance=0.1):
    """plot_accuracy_keyframe.m:41-53,150-165 transcribed step by step (1-based index arithmetic kept)."""
    D = np.array(D, dtype=np.float64)
    D[D > max_distance] = np.inf
    d = np.sort(D)
    n = d.size
    accuracy = np.cumsum(np.ones(n)) / n
    index = np.isfinite(d)
    rec, prec = d[index], accuracy[index]
    mrec = np.concatenate([[0.0], rec, [0.1]])
    mpre = np.concatenate([[0.0], prec, [prec[-1]]])
    for i in range(1, mpre.size):
        mpre[i] = max(mpre[i], mpre[i - 1])
    ap = 0.0
    for i in range(1, mrec.size):
        if mrec[i] != mrec[i - 1]:
            ap += (mrec[i] - mrec[i - 1]) * mpre[i]
    return ap * 10.0


def test_auc_matches_matlab_transcription():
    from densefusion_b200 import evaluation as E
    rng = np.random.default_rng(1)
    for n in (1, 7, 200):
        d = np.abs(rng.normal(0.03, 0.04, size=n))
        d[::5] = d[0]                                           # ties
        assert abs(E.auc(d) - _vocap_matlab(d)) < 1e-12
    assert E.auc(np.zeros(10)) == pytest.approx(1.0)
    assert E.auc(np.full(10, 0.5)) == 0.0                       # everything beyond 10 cm


def test_auc_hand_computed_vectors():
    """VOCap (replace_ycb_toolbox/plot_accuracy_keyframe.m:150-170) worked by hand, independent of any transcription:

    d = [0.02 0.04 0.04 0.06 0.2], max_distance 0.1 -> 0.2 becomes inf; accuracy = [.2 .4 .6 .8 1]; finite part:
      mrec = [0 .02 .04 .04 .06 .1], mpre = [0 .2 .4 .6 .8 .8] (already monotone);
      steps where mrec changes: .02*.2 + .02*.4 + (.04->.04: none) + .02*.8 + .04*.8 = .004 + .008 + .016 + .032 = .06; x10 = 0.6
    d = [0.05]:            mrec = [0 .05 .1], mpre = [0 1 1] -> .05*1 + .05*1 = .1 -> 1.0
    d = [0.03 0.03 0.5 0.5]: accuracy = [.25 .5 .75 1], finite mrec = [0 .03 .03 .1], mpre = [0 .25 .5 .5]
                           -> .03*.25 + .07*.5 = .0075 + .035 = .0425 -> 0.425
    d = [0.1 0.1]:         0.1 is not > max_distance: mrec = [0 .1 .1 .1], mpre = [0 .5 1 1] -> .1*.5 = .05 -> 0.5"""
    from densefusion_b200 import evaluation as E
    assert E.auc([0.02, 0.04, 0.04, 0.06, 0.2]) == pytest.approx(0.6, abs=1e-12)
    assert E.auc([0.06, 0.2, 0.04, 0.02, 0.04]) == pytest.approx(0.6, abs=1e-12)      # order of the inputs is irrelevant
    assert E.auc([0.05]) == pytest.approx(1.0, abs=1e-12)
    assert E.auc([0.03, 0.03, 0.5, 0.5]) == pytest.approx(0.425, abs=1e-12)
    assert E.auc([0.1, 0.1]) == pytest.approx(0.5, abs=1e-12)


@pytest.mark.gpu
def test_pose_distances_on_reference_point_clouds():
    """The root .ply pair of the reference tree (tools/eval_cad.py:130-136 output): ADD 0.0168566, ADD-S 0.0092865
     -- identity pose, model = predicted cloud, target = ground-truth cloud."""
    from densefusion_b200 import evaluation as E
    g = golden("ply_pair")
    pred = torch.from_numpy(g["pred"].astype(np.float32)).cuda()[None]
    tgt = torch.from_numpy(g["target"].astype(np.float32)).cuda()[None]
    pose = torch.tensor([[1.0, 0, 0, 0, 0, 0, 0]], dtype=torch.float64, device="cuda")
    add = E.pose_distances(pose, pred, tgt, torch.tensor([3], device="cuda"), [12])
    adds = E.pose_distances(pose, pred, tgt, torch.tensor([12], device="cuda"), [12])
    assert abs(float(add) - float(g["add"])) < 1e-6 and abs(float(adds) - float(g["adds"])) < 1e-6
    per_obj, overall = E.success_rates(torch.cat([add, adds]), [3, 12], [0.01] * 21, 21)
    assert per_obj[3] == 0.0 and per_obj[12] == 1.0 and overall == 0.5


@pytest.mark.gpu
def test_add_adi_vs_float64_numpy():
    from densefusion_b200 import evaluation as E
    from densefusion_b200 import synth
    g = torch.Generator().manual_seed(3)
    pts = torch.randn(300, 3, generator=g) * 0.05
    q = torch.stack([synth.random_unit_quaternion(g) for _ in range(4)]).double()
    qg = torch.stack([synth.random_unit_quaternion(g) for _ in range(4)]).double()
    t, tg = torch.randn(4, 3, generator=g).double() * 0.1, torch.randn(4, 3, generator=g).double() * 0.1
    poses, poses_gt = torch.cat([q, t], 1).cuda(), torch.cat([qg, tg], 1).cuda()
    add = E.add_distances(poses, poses_gt, pts.cuda()).cpu().numpy()
    adi = E.adi_distances(poses, poses_gt, pts.cuda()).cpu().numpy()
    for b in range(4):
        Re, Rg = synth.quat_to_rot(q[b]).double().numpy(), synth.quat_to_rot(qg[b]).double().numpy()
        pe = pts.double().numpy() @ Re.T + t[b].numpy()
        pg = pts.double().numpy() @ Rg.T + tg[b].numpy()
        assert abs(add[b] - np.mean(np.linalg.norm(pe - pg, axis=1))) < 1e-6
        dm = np.linalg.norm(pg[:, None] - pe[None], axis=2).min(1)
        assert abs(adi[b] - dm.mean()) < 1e-6
