"""CPU-only checks around the data-parallel training step (config C4):
  * the oracle's gradient / Adam restatement pinned against the reference itself (tests/golden/c4_train_ycb.npz,
    tools/train.py:143-169 run on the reference modules by tests/golden/make_golden.py);
  * host logic of densefusion_b200.trainer: the flat arena aliases parameters and gradients, autograd accumulates
    into it, the world_size-2 gloo all-reduce equals serial accumulation, shard_range partitions."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden
from densefusion_b200 import synth
from oracle import df_oracle as O


def _summary_close(g, prefix, grads, rtol=2e-4):
    worst = 0.0
    for name, grad in grads.items():
        if prefix + "gnone." + name in g.files:
            assert grad is None or float(grad.abs().max()) == 0.0, name
            continue
        f = grad.reshape(-1).double()
        stride = max(1, f.numel() // 32)
        want_norm = float(g[prefix + "gnorm." + name])
        samp = g[prefix + "gsamp." + name].astype(np.float64)
        got = f[::stride][:32].numpy()
        scale = max(want_norm / max(f.numel(), 1) ** 0.5, 1e-12)          # rms of the tensor
        err = max(abs(float(f.norm()) - want_norm) / max(want_norm, 1e-30), float(np.max(np.abs(got - samp))) / scale / 30)
        worst = max(worst, err)
        assert abs(float(f.norm()) - want_norm) <= rtol * max(want_norm, 1e-12), (name, float(f.norm()), want_norm)
        assert np.max(np.abs(got - samp)) <= 30 * rtol * scale + 1e-9, name
    return worst


def _crops(g):
    n, o, m, h, w, seed, iters = [int(v) for v in g["meta"]]
    crops = [synth.synth_crop(int(c), n, m, o, (h, w), int(ob)) for c, ob in zip(g["cases"], g["objs"])]
    return crops, n, o, m, seed, iters


def test_oracle_training_gradients_vs_reference_golden():
    g = golden("c4_train_ycb")
    crops, n, o, m, seed, iters = _crops(g)
    from densefusion_b200.lib.network import PoseNet, PoseRefineNet
    est_sd = synth.synth_state_dict(synth.shapes_of(PoseNet(n, o)), seed)
    ref_sd = synth.synth_state_dict(synth.shapes_of(PoseRefineNet(n, o)), seed + 1)
    sym, w = [int(s) for s in g["sym_list"]], float(g["w"])
    grads, losses, dists = O.estimator_gradients(est_sd, crops, o, m, sym, w)
    assert np.allclose(losses, g["est_losses"], rtol=1e-5) and np.allclose(dists, g["est_dis"], rtol=1e-5)
    print("estimator grad worst err", _summary_close(g, "est.", grads))
    # one Adam step of the reference's optimiser
    params = {k: v.clone() for k, v in est_sd.items()}
    O.adam_reference(params, grads, {}, lr=1e-4)
    for name in params:
        f = (params[name] - est_sd[name]).reshape(-1)
        got = f[::max(1, f.numel() // 32)][:32].numpy()
        want = g["est.delta." + name]
        # step-1 Adam moves every touched weight by ~lr*sign(g): compare where the gradient is not vanishing
        assert np.mean(np.abs(got - want) < 2e-6) > 0.95, name
    rgrads, rd = O.refiner_gradients(est_sd, ref_sd, crops, o, m, sym, w, iters)
    assert np.allclose(rd, g["ref_dis"], rtol=1e-5)
    print("refiner grad worst err", _summary_close(g, "ref.", rgrads))


def test_flat_arena_aliases_and_accumulates():
    from densefusion_b200.trainer import FlatArena
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    before = [p.detach().clone() for p in net.parameters()]
    arena = FlatArena(net.parameters())
    assert arena.numel == sum(p.numel() for p in net.parameters()) and arena.total % 64 == 0
    for p, b, o in zip(net.parameters(), before, arena.offsets):
        assert torch.equal(p.detach(), b) and o % 64 == 0
        assert p.data_ptr() == arena.param.data_ptr() + 4 * o and p.grad.data_ptr() == arena.grad.data_ptr() + 4 * o
    x = torch.randn(4, 5)
    net(x).sum().backward()
    net(2 * x).sum().backward()                      # accumulates in place, like tools/train.py:159-169
    g1 = arena.grad.clone()
    assert float(g1.abs().sum()) > 0
    ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    ref.load_state_dict(net.state_dict())
    (ref(x).sum() + ref(2 * x).sum()).backward()
    for p, q in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-6)
    arena.zero_grad()
    assert float(arena.grad.abs().sum()) == 0.0 and all(float(p.grad.abs().sum()) == 0.0 for p in net.parameters())
    with pytest.raises(Exception):
        arena.adam_step(1e-4)                        # the optimiser step is a CUDA kernel: no CPU fallback


def test_shard_range_partitions():
    from densefusion_b200.trainer import shard_range
    for n in (0, 1, 7, 8, 100, 257):
        for world in (1, 2, 3, 8):
            got = [i for r in range(world) for i in shard_range(n, r, world)]
            assert got == list(range(n))
            sizes = [len(shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from densefusion_b200.trainer import FlatArena, shard_range
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 9), torch.nn.ReLU(), torch.nn.Linear(9, 2))
        arena = FlatArena(net.parameters())
        data = torch.randn(10, 6, generator=torch.Generator().manual_seed(1))
        arena.zero_grad()
        for i in shard_range(10, rank, world):           # every rank: its shard, one backward per sample
            net(data[i:i + 1]).pow(2).sum().backward()
        arena.all_reduce()
        ret[rank] = arena.grad.clone()
    finally:
        dist.destroy_process_group()


def test_gloo_world2_allreduce_equals_serial_accumulation():
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_dp_worker, args=(world, port, ret), nprocs=world, join=True)
    from densefusion_b200.trainer import FlatArena
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 9), torch.nn.ReLU(), torch.nn.Linear(9, 2))
    arena = FlatArena(net.parameters())
    data = torch.randn(10, 6, generator=torch.Generator().manual_seed(1))
    for i in range(10):
        net(data[i:i + 1]).pow(2).sum().backward()
    assert torch.allclose(ret[0], arena.grad, atol=1e-5) and torch.equal(ret[0], ret[1])


def _dp_overlap_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from densefusion_b200.trainer import FlatArena, shard_range
        torch.manual_seed(0)
        # `unused` never receives a gradient (like the encoder's classifier): its bucket is exchanged by the final flush
        net = torch.nn.ModuleDict({"a": torch.nn.Linear(6, 40), "b": torch.nn.Linear(40, 40), "c": torch.nn.Linear(40, 2),
                                   "unused": torch.nn.Linear(3, 3)})
        arena = FlatArena(net.parameters())
        arena.enable_overlap(bucket_bytes=4 * 300)          # several buckets
        assert len(arena._buckets) >= 3 and arena._buckets[-1][0] == 0 and arena._buckets[0][1] == arena.total
        data = torch.randn(12, 6, generator=torch.Generator().manual_seed(1))
        mine = list(shard_range(12, rank, world))
        launched_early = []
        for step in range(2):                                # two optimiser steps: the counters re-arm
            arena.zero_grad()
            assert arena.begin_step(expected=len(mine))
            for i in mine:
                net["c"](torch.relu(net["b"](torch.relu(net["a"](data[i:i + 1]))))).pow(2).sum().backward()
            launched_early.append(sum(arena._launched))
            arena.finish_all_reduce()
        ret[rank] = (arena.grad.clone(), launched_early, len(arena._buckets))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_overlapped_bucketed_allreduce_equals_serial_accumulation():
    """The bucketed exchange launched from the parameters' post-accumulate hooks (trainer.FlatArena.enable_overlap) gives the same
    summed gradients as serial accumulation, launches the gradient-carrying buckets before the backward pass has ended, and
    leaves parameters that never receive a gradient to the final flush."""
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_dp_overlap_worker, args=(world, port, ret), nprocs=world, join=True)
    from densefusion_b200.trainer import FlatArena
    torch.manual_seed(0)
    net = torch.nn.ModuleDict({"a": torch.nn.Linear(6, 40), "b": torch.nn.Linear(40, 40), "c": torch.nn.Linear(40, 2),
                               "unused": torch.nn.Linear(3, 3)})
    arena = FlatArena(net.parameters())
    data = torch.randn(12, 6, generator=torch.Generator().manual_seed(1))
    for i in range(12):
        net["c"](torch.relu(net["b"](torch.relu(net["a"](data[i:i + 1]))))).pow(2).sum().backward()
    g0, early, nb = ret[0]
    assert torch.allclose(g0, arena.grad, atol=1e-5) and torch.equal(g0, ret[1][0])
    # first step: the bucket that also holds `unused` waits for the final flush; afterwards the arena knows which parameters
    # receive gradients and every bucket goes out from the hooks, before the backward pass has ended
    assert 0 < early[0] < nb and early[1] == nb
