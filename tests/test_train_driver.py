"""Training-driver semantics: the schedule state machine of tools/train.py:211-251 on CPU, and a
tiny end-to-end run of the driver on the GPU (phase switch, checkpoint names and keys, resume)."""
import os

import pytest
import torch


def test_schedule_follows_reference_state_machine():
    from densefusion_b200.train_driver import Schedule, TrainOptions
    s = Schedule(TrainOptions(batch_size=8, lr=1e-4, lr_rate=0.3, w=0.015, w_rate=0.3, decay_margin=0.016,
                              refine_margin=0.013, iteration=2))
    assert s.phase == "estimator" and s.batch_size == 8
    a = s.after_test(1, 0.05)
    assert a == [("save_best", "pose_model_1_0.05.pth")]
    assert s.after_test(2, 0.06) == []                                  # not a new best: nothing happens
    a = s.after_test(3, 0.0155)                                         # below decay_margin only
    assert [k for k, _ in a] == ["save_best", "decay"] and s.decay_start and not s.refine_start
    assert s.lr == pytest.approx(3e-5) and s.w == pytest.approx(0.0045)
    assert s.after_test(4, 0.0150)[0] == ("save_best", "pose_model_4_0.015.pth")      # decay happens once
    assert s.lr == pytest.approx(3e-5)
    a = s.after_test(5, 0.0129)                                         # below refine_margin: the best estimator is saved first
    assert [k for k, _ in a] == ["save_best", "refine_start"] and a[0][1] == "pose_model_5_0.0129.pth"
    assert s.phase == "refiner" and s.batch_size == 4
    assert s.current_name() == "pose_refine_model_current.pth"
    assert s.after_test(6, 0.0100) == [("save_best", "pose_refine_model_6_0.01.pth")]
    # one epoch that crosses both margins at once: decay and refine start in the same epoch, in the reference's order
    s2 = Schedule(TrainOptions())
    assert [k for k, _ in s2.after_test(1, 0.010)] == ["save_best", "decay", "refine_start"]
    # resuming a refiner checkpoint starts decayed, in the refiner phase, with the halved batch (tools/train.py:86-93)
    s3 = Schedule(TrainOptions(resume_refinenet="pose_refine_model_current.pth"))
    assert s3.phase == "refiner" and s3.decay_start and s3.batch_size == 4 and s3.lr == pytest.approx(3e-5)


def test_synthetic_dataset_contract():
    from densefusion_b200.train_driver import SyntheticPoseDataset, collate_buckets
    ds = SyntheticPoseDataset("train", 500, 7, refine=False)
    pts, choose, img, target, model, idx = ds[0]
    assert pts.shape == (500, 3) and choose.shape == (1, 500) and img.shape[0] == 3 and target.shape == (500, 3)
    assert model.shape == (500, 3) and idx.shape == (1,) and choose.dtype == torch.int64
    assert SyntheticPoseDataset("train", 500, 7, refine=True).get_num_points_mesh() == 2600
    buckets = collate_buckets([ds[i] for i in range(6)], "cpu")
    assert sorted(b["img"].shape[0] for b in buckets) == [2, 2, 2] and all(b["idx"].shape[1] == 1 for b in buckets)


@pytest.mark.gpu
def test_driver_runs_both_phases_and_writes_reference_checkpoints(tmp_path):
    from densefusion_b200.lib.network import PoseNet, PoseRefineNet
    from densefusion_b200.train_driver import SyntheticPoseDataset, TrainDriver, TrainOptions
    torch.backends.cudnn.allow_tf32 = False
    est, ref = PoseNet(500, 21).cuda(), PoseRefineNet(500, 21).cuda()
    opt = TrainOptions(batch_size=4, iteration=2, outf=str(tmp_path), save_every=8,
                       decay_margin=10.0, refine_margin=5.0)            # random weights give dis ~ 0.5: both switches fire
    factory = lambda mode, refine: SyntheticPoseDataset(mode, 500, 8 if mode == "train" else 4, refine=refine,
                                                        sizes=((80, 80),))
    drv = TrainDriver(est, ref, factory, opt, log=lambda s: None)
    before = {k: v.clone() for k, v in est.state_dict().items()}
    hist = drv.run(epochs=2)
    assert hist[0]["phase"] == "refiner" and hist[0]["actions"] == ["save_best", "decay", "refine_start"]
    assert hist[0]["batch_size"] == 2 and hist[0]["lr"] == pytest.approx(3e-5)
    files = sorted(os.listdir(tmp_path))
    assert "pose_model_current.pth" in files and any(f.startswith("pose_model_1_") for f in files)
    assert any(f.startswith("pose_refine_model_") for f in files)
    assert drv.num_points_mesh == 2600 and drv.trainer.phase == "refiner"
    sd = torch.load(os.path.join(tmp_path, [f for f in files if f.startswith("pose_model_1_")][0]))
    assert set(sd) == set(before) and "cnn.model.module.feats.conv1.weight" in sd
    assert any(not torch.equal(sd[k].cpu(), before[k].cpu()) for k in sd)                  # the estimator phase trained it
    est_after = {k: v.clone() for k, v in est.state_dict().items()}
    drv.run(epochs=1)                                                   # refiner phase leaves the estimator untouched
    assert all(torch.equal(est.state_dict()[k], est_after[k]) for k in est_after)
    # resume into the refiner phase
    opt2 = TrainOptions(batch_size=4, outf=str(tmp_path), resume_posenet="pose_model_current.pth",
                        resume_refinenet=[f for f in files if f.startswith("pose_refine_model_")][0])
    drv2 = TrainDriver(PoseNet(500, 21).cuda(), PoseRefineNet(500, 21).cuda(), factory, opt2, log=lambda s: None)
    assert drv2.sched.phase == "refiner" and drv2.trainer.phase == "refiner" and drv2.sched.batch_size == 2
