"""K4 parity: indices bit-exact against (i) the C emulator oracle/knn_ref.c and (ii) the reference's own
CUDA kernels compiled from /root/reference (oracle/_ref), incl. ties, duplicates, NaN/Inf, ragged sizes."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import df_oracle as O
from util import reference_knn_gpu

pytestmark = pytest.mark.gpu


def _ours(ref, qry, k=1):
    from densefusion_b200.lib.knn import KNearestNeighbor
    return KNearestNeighbor(k)(ref, qry).cpu()


def _check(ref, qry, k=1, use_ref_kernel=True):
    ours = _ours(ref, qry, k)
    want = O.knn(ref, qry, k)
    assert ours.dtype == torch.int64 and tuple(ours.shape) == (ref.shape[0], k, qry.shape[2])
    assert torch.equal(ours, want), f"mismatch vs C emulator: {(ours != want).sum().item()} of {want.numel()}"
    if use_ref_kernel:
        for b in range(ref.shape[0]):
            r = reference_knn_gpu(ref[b].cuda().contiguous(), qry[b].cuda().contiguous(), k)
            if r is not None:
                assert torch.equal(ours[b], r.cpu()), "mismatch vs the reference CUDA kernel"


@pytest.mark.parametrize("R,Q", [(500, 250000), (37, 53), (1, 7), (2049, 1025), (16, 256), (17, 257), (4100, 3)])
def test_knn1_random(R, Q):
    g = torch.Generator().manual_seed(R * 7919 + Q)
    _check(torch.randn(1, 3, R, generator=g) * 0.05, torch.randn(1, 3, Q, generator=g) * 0.05)


def test_knn1_batched():
    g = torch.Generator().manual_seed(3)
    _check(torch.randn(3, 3, 300, generator=g), torch.randn(3, 3, 1111, generator=g))


# Q <= 18 944 takes the warp-per-query kernel (lanes split the references, lexicographic (distance, index) shuffle reduction),
# larger Q the thread-per-query kernel: the tie / duplicate / NaN sets run through both
@pytest.mark.parametrize("Q", [5000, 40000])
def test_knn1_lattice_ties_and_duplicates(Q):
    g = torch.Generator().manual_seed(11)
    ref = torch.randint(0, 8, (1, 3, 700), generator=g).float() / 1024.0       # many exact ties
    ref[0, :, 350:] = ref[0, :, :350]                                          # exact duplicates
    qry = torch.randint(0, 8, (1, 3, Q), generator=g).float() / 1024.0
    _check(ref, qry)
    ours = _ours(ref, qry).view(-1)
    assert int(ours.max()) <= 350                                               # lowest index of each duplicate pair


@pytest.mark.parametrize("Q", [999, 25000])
def test_knn1_nan_inf_rows(Q):
    g = torch.Generator().manual_seed(12)
    ref = torch.randn(1, 3, 100, generator=g)
    qry = torch.randn(1, 3, Q, generator=g)
    ref[0, 0, 5] = float("nan")
    ref[0, 1, 17] = float("inf")
    qry[0, 2, 3] = float("nan")
    qry[0, 0, 4] = float("inf")
    _check(ref, qry)
    ref[0, 0, 0] = float("nan")              # NaN in row 0: the reference never leaves index 1
    _check(ref, qry)
    assert torch.all(_ours(ref, qry) == 1)
    # every distance +inf (an infinite coordinate in every reference): the seed row stays; and +inf rows beside finite ones
    ref = torch.randn(1, 3, 100, generator=g)
    ref[0, 0, :] = float("inf")
    _check(ref, qry)
    assert torch.all(_ours(ref, qry) == 1)
    ref[0, 0, 40:] = 0.25
    _check(ref, qry)


@pytest.mark.parametrize("R", [500, 2600])
def test_knn1_small_q_shapes_of_the_refiner_loss(R):
    """Q = R = num_pt_mesh (lib/loss_refiner.py:40-46, tools/eval_linemod.py:124-128): the warp-per-query path, incl. the model
    queried against itself (every point is its own nearest neighbour or an earlier duplicate's)."""
    g = torch.Generator().manual_seed(R)
    pts = torch.randn(1, 3, R, generator=g) * 0.05
    pts[0, :, R // 2:R // 2 + 20] = pts[0, :, 0:20]                            # duplicates: the earlier index must win
    _check(pts, pts)
    _check(pts, pts + torch.randn(1, 3, R, generator=g) * 1e-3)
    self_ind = _ours(pts, pts).view(-1)
    assert torch.all(self_ind <= torch.arange(1, R + 1))


@pytest.mark.parametrize("D,R,Q,k", [(128, 100, 1000, 2), (3, 64, 500, 5), (5, 33, 70, 1), (3, 40, 40, 40)])
def test_knn_general(D, R, Q, k):
    g = torch.Generator().manual_seed(D + R + Q + k)
    _check(torch.rand(2, D, R, generator=g), torch.rand(2, D, Q, generator=g), k)


def test_knn_ply_fixture():
    gz = golden("ply_pair")
    p32 = torch.from_numpy(gz["pred"].astype(np.float32)).t().contiguous()[None]
    t32 = torch.from_numpy(gz["target"].astype(np.float32)).t().contiguous()[None]
    inds = _ours(t32, p32).view(-1) - 1
    assert np.array_equal(inds.numpy().astype(np.int16), gz["inds"])


def test_knn_errors():
    from densefusion_b200 import _C, ops
    with pytest.raises(_C.DFError):
        ops.knn(torch.zeros(1, 3, 4, device="cuda"), torch.zeros(2, 3, 4, device="cuda"))
    with pytest.raises(_C.DFError):
        ops.knn(torch.zeros(3, 4, device="cuda"), torch.zeros(1, 3, 4, device="cuda"))
    with pytest.raises(_C.DFError):
        ops.knn(torch.zeros(1, 3, 4), torch.zeros(1, 3, 4))                     # CPU tensors: no CPU path


@pytest.mark.parametrize("R,P", [(500, 4096), (2600, 2048), (20000, 500)])
def test_knn_sweep_full_size_properties(R, P):
    """Config C3 at full size: Q = P*R queries.  Size-independent checks: (a) every reference point queried
    against the set returns itself or an earlier duplicate, (b) a random 100k subset equals the reference
    CUDA kernel bit for bit, (c) the distance at the returned index equals the brute-force minimum."""
    from densefusion_b200 import ops
    g = torch.Generator().manual_seed(R + P)
    ref = (torch.randn(1, 3, R, generator=g) * 0.05).cuda()
    Q = min(P * R, 20_000_000)
    qry = (torch.randn(1, 3, Q, generator=g) * 0.06).cuda()
    ind = ops.knn(ref, qry, 1).view(-1)
    assert int(ind.min()) >= 1 and int(ind.max()) <= R
    self_ind = ops.knn(ref, ref, 1).view(-1)
    assert torch.all(self_ind <= torch.arange(1, R + 1, device="cuda"))
    d_self = ((ref[0] - ref[0][:, self_ind - 1]) ** 2).sum(0)
    assert torch.all(d_self == 0)
    sub = torch.randperm(Q, generator=g)[:100_000].cuda()
    qs = qry[0][:, sub].contiguous()
    r = reference_knn_gpu(ref[0].contiguous(), qs, 1)
    if r is not None:
        assert torch.equal(ind[sub], r.view(-1))
    d = ((ref[0][:, :, None] - qs[:, None, :4096]) ** 2).sum(0)                # (R, 4096)
    got = torch.gather(d, 0, (ind[sub][:4096] - 1)[None]).view(-1)
    assert torch.allclose(got, d.min(0)[0], rtol=1e-5, atol=1e-12)
