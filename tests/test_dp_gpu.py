"""Data-parallel training step over NCCL (config C4): two ranks, each with its shard of the batch, must end the step
with the same parameters as one rank that saw the whole batch (gradient SUM semantics, tools/train.py:159-169).
Needs 2 GPUs (run with `gpurun --gpus 2`); skipped on a single-GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from densefusion_b200 import synth

pytestmark = pytest.mark.gpu


def _batch(cases, objs, dev):
    crops = [synth.synth_crop(c, 500, 500, 21, (80, 80), o) for c, o in zip(cases, objs)]
    keys = ("img", "points", "choose", "idx", "target", "model_points")
    return {k: torch.cat([c[k] for c in crops], 0).to(dev) for k in keys}


CASES, OBJS = [60, 61, 62, 63], [12, 3, 7, 15]


def _train(dev, cases, objs, phase, group_ok):
    from densefusion_b200.lib.network import PoseNet, PoseRefineNet
    from densefusion_b200.trainer import DataParallelTrainer
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    est, ref = PoseNet(500, 21), PoseRefineNet(500, 21)
    est.load_state_dict(synth.synth_state_dict(synth.shapes_of(est), 8))
    ref.load_state_dict(synth.synth_state_dict(synth.shapes_of(ref), 9))
    est.eval().to(dev)
    ref.eval().to(dev)
    tr = DataParallelTrainer(est, ref, 500, synth.YCB_SYM, lr=1e-4, iteration=2, phase=phase)
    for _ in range(2):
        tr.step([_batch(cases, objs, dev)])
    arena = tr.arena_est if phase == "estimator" else tr.arena_ref
    return arena.param.detach().clone()


def _worker(rank, world, port, phase, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from densefusion_b200.trainer import shard_range
        mine = list(shard_range(len(CASES), rank, world))
        p = _train(dev, [CASES[i] for i in mine], [OBJS[i] for i in mine], phase, True)
        ret[rank] = p.cpu()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("phase", ["estimator", "refiner"])
def test_two_rank_step_equals_single_rank_step(phase):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, phase, ret), nprocs=2, join=True)
    assert torch.equal(ret[0], ret[1]), "ranks diverged after the all-reduce + Adam"
    single = _train(torch.device("cuda", 0), CASES, OBJS, phase, False).cpu()
    d = (ret[0] - single).abs()
    # two Adam steps of lr 1e-4: identical up to the summation order of the per-rank partial gradients
    share = float((d > 2e-6).float().mean())
    print(f"{phase}: max |dp - single| = {float(d.max()):.3e}, share of weights off by > 2e-6: {share:.5f}")
    assert share < 0.01
