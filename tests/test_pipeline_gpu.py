"""K5 + subsystem (4): on-device pose selection / composition and the estimate + refine pipeline against the
oracle's float64 host algebra and the reference's eval-loop outputs (tests/golden)."""
import numpy as np
import pytest
import torch

from conftest import golden
from densefusion_b200 import synth
from oracle import df_oracle as O
from util import build_nets, rel

pytestmark = pytest.mark.gpu


def test_select_transform_compose_vs_numpy():
    from densefusion_b200 import ops
    B, n = 6, 500
    g = torch.Generator().manual_seed(17)
    pred_r = torch.randn(B, n, 4, generator=g)
    pred_t = torch.randn(B, n, 3, generator=g) * 0.02
    pred_c = torch.rand(B, n, 1, generator=g)
    pred_c[2, 10, 0] = 2.0
    pred_c[2, 300, 0] = 2.0                                     # tie: first index wins
    cloud = torch.randn(B, n, 3, generator=g) * 0.05 + torch.tensor([0.0, 0.0, 0.8])
    pose, which = ops.select_pose(pred_r.cuda(), pred_t.cuda(), pred_c.cuda(), cloud.cuda())
    assert int(which[2]) == 10
    r2 = torch.randn(B, 4, generator=g)
    # exercise every pivot branch of quaternion_from_matrix: rotations by ~pi about x / y / z, and identity-ish
    r2[0] = torch.tensor([1e-3, 1.0, 0.02, 0.01]); r2[1] = torch.tensor([1e-3, 0.02, 1.0, 0.01])
    r2[3] = torch.tensor([1e-3, 0.01, 0.02, 1.0]); r2[4] = torch.tensor([1.0, 1e-4, 0.0, 0.0])
    t2 = torch.randn(B, 3, generator=g) * 0.01
    new_cloud = ops.cloud_transform(cloud.cuda(), pose)
    pose0 = pose.clone()
    ops.pose_compose_(pose, r2.cuda(), t2.cuda())
    for b in range(B):
        my_r, my_t, wm = O.select_pose(pred_r[b:b + 1], pred_t[b:b + 1], pred_c[b:b + 1], cloud[b:b + 1])
        assert wm == int(which[b])
        assert np.allclose(pose0[b].cpu().numpy(), np.append(my_r, my_t).astype(np.float64), rtol=0, atol=1e-7)
        m1 = O.quaternion_matrix(my_r)
        R = torch.from_numpy(m1[:3, :3].astype(np.float32))
        T = torch.from_numpy(my_t.astype(np.float32))
        want_cloud = (cloud[b] - T) @ R
        assert rel(new_cloud[b], want_cloud) < 1e-6
        m1[0:3, 3] = my_t
        q2 = (r2[b] / r2[b].norm()).numpy()
        m2 = O.quaternion_matrix(q2)
        m2[0:3, 3] = t2[b].numpy()
        final = m1 @ m2
        rot = final.copy(); rot[0:3, 3] = 0
        want = np.append(O.quaternion_from_matrix(rot, True), final[0:3, 3])
        assert np.allclose(pose[b].cpu().numpy(), want, rtol=0, atol=1e-6), (b, pose[b].cpu().numpy(), want)


@pytest.mark.parametrize("name", ["c0_linemod_add", "c1_ycb_adds"])
def test_estimate_refine_vs_reference_golden(name):
    from densefusion_b200.pipeline import PoseEstimator
    g = golden(name)
    case, n, o, m, h, w, obj, seed, iters = [int(v) for v in g["meta"]]
    est, ref, _, _ = build_nets(n, o, seed)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    d = {k: v.cuda() for k, v in synth.synth_crop(case, n, m, o, (h, w), obj).items()}
    pipe = PoseEstimator(est, ref, iterations=iters, precision="fp32")
    for n_it, key in ((0, "pose0"), (1, "pose_iter1"), (iters, f"pose_iter{iters}"), (4, "pose_iter4")):
        pose = pipe.estimate(d["img"], d["points"], d["choose"], d["idx"], iterations=n_it).cpu().numpy()[0]
        assert rel(pose, g[key]) < 1e-4, (key, pose, g[key])
    # head + refine on the reference's own embedding (isolates the encoder)
    emb_pm = torch.from_numpy(g["emb"][0].T.copy()).cuda()
    pose = pipe.head_and_refine(d["points"], emb_pm, d["idx"], iterations=iters).cpu().numpy()[0]
    assert rel(pose, g[f"pose_iter{iters}"]) < 1e-4


def test_batched_buckets_graph_and_oracle():
    """8-object frames with mixed crop sizes: bucketed batch == per-crop results == oracle; graph replay == eager."""
    from densefusion_b200.pipeline import GraphedBuckets, PoseEstimator
    n, o, m = 500, 21, 500
    est, ref, est_sd, ref_sd = build_nets(n, o, seed=8)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sizes = [(80, 80), (120, 120)]
    buckets_cpu = [synth.batch_crops(range(50 + 10 * i, 53 + 10 * i), num_points=n, num_pt_mesh=m, num_obj=o, hw=hw)
                   for i, hw in enumerate(sizes)]
    buckets = [dict(img=b["img"].cuda(), cloud=b["points"].cuda(), choose=b["choose"].cuda(), obj=b["idx"].view(-1).cuda())
               for b in buckets_cpu]
    pipe = PoseEstimator(est, ref, iterations=2, precision="fp32", chunk_crops=4)       # 6 crops -> 2 chunks
    poses = pipe.estimate_buckets(buckets).cpu().numpy().copy()
    assert poses.shape == (6, 7)
    k = 0
    for b in buckets_cpu:
        for i in range(b["points"].shape[0]):
            want = O.estimate_and_refine(est_sd, ref_sd, b["img"][i:i + 1], b["points"][i:i + 1], b["choose"][i:i + 1],
                                         b["idx"][i:i + 1], o, 2)
            assert rel(poses[k], want) < 1e-4, (k, poses[k], want)
            k += 1
    graphed = GraphedBuckets(pipe, [(3, 80, 80), (3, 120, 120)])
    host = [dict(img=b["img"].pin_memory(), cloud=b["points"].pin_memory(), choose=b["choose"].pin_memory(),
                 obj=b["idx"].view(-1).pin_memory()) for b in buckets_cpu]
    graphed.load(host)
    out = graphed.run().cpu().numpy()
    assert np.array_equal(out, poses)
    out2 = graphed.run().cpu().numpy()
    assert np.array_equal(out2, poses)


@pytest.mark.parametrize("n,precision", [(500, "hybrid16"), (1000, "hybrid16"), (500, "hybrid"), (500, "hybrid16s"), (1000, "hybrid16s")])
def test_mixed_buckets_graphed_tensor_core_vs_oracle(n, precision):
    """The bench's configuration in small: three crop-size buckets, the default tensor-core arithmetic, chunked head, side
    streams per bucket, CUDA-graph replay -- every pose against the oracle's estimate + 2 refine iterations (<= 1e-4 in the
    max-norm of the 7-vector and element-wise with an absolute floor), also at the reference's N = 1000 points."""
    from densefusion_b200.pipeline import GraphedBuckets, PoseEstimator
    from util import rel_elementwise
    o, m = 21, 500
    est, ref, est_sd, ref_sd = build_nets(n, o, seed=11)
    sizes = [(80, 80), (120, 120), (160, 160)]
    counts = [3, 2, 2]
    buckets_cpu = [synth.batch_crops(range(70 + 10 * i, 70 + 10 * i + c), num_points=n, num_pt_mesh=m, num_obj=o, hw=hw)
                   for i, (hw, c) in enumerate(zip(sizes, counts))]
    pipe = PoseEstimator(est, ref, iterations=2, precision=precision, chunk_crops=4)       # 7 crops -> 2 chunks
    graphed = GraphedBuckets(pipe, [(c, h, w) for c, (h, w) in zip(counts, sizes)])
    host = [dict(img=b["img"].pin_memory(), cloud=b["points"].pin_memory(), choose=b["choose"].pin_memory(),
                 obj=b["idx"].view(-1).pin_memory()) for b in buckets_cpu]
    graphed.load(host)
    poses = graphed.run().cpu().numpy().copy()
    assert poses.shape == (sum(counts), 7)
    k, worst = 0, 0.0
    for b in buckets_cpu:
        for i in range(b["points"].shape[0]):
            want = O.estimate_and_refine(est_sd, ref_sd, b["img"][i:i + 1], b["points"][i:i + 1], b["choose"][i:i + 1],
                                         b["idx"][i:i + 1], o, 2)
            e, ee = rel(poses[k], want), rel_elementwise(poses[k], want, floor=1e-2)
            worst = max(worst, e)
            assert e < 1e-4 and ee < 2e-3, (k, e, ee, poses[k], want)
            k += 1
    print(f"graphed mixed buckets {precision} n={n}: worst pose error {worst:.3e}")


def test_streaming_estimator_equals_direct_calls():
    """Double-buffered serving loop (H2D of batch i+1 overlapping the compute of batch i) returns, for every batch,
    exactly the poses of a direct estimate_buckets call."""
    import torch
    from densefusion_b200.pipeline import PoseEstimator, StreamingEstimator
    est, ref, _, _ = build_nets(500, 21, seed=0)
    pipe = PoseEstimator(est, ref, iterations=2, precision="3xtf32")
    shapes = [(2, 80, 80), (1, 120, 120)]

    def host_batch(seed):
        g = torch.Generator().manual_seed(seed)
        out = []
        for b, h, w in shapes:
            out.append({"img": torch.randn(b, 3, h, w, generator=g).pin_memory(),
                        "cloud": (torch.randn(b, 500, 3, generator=g) * 0.05 + torch.tensor([0.0, 0.0, 0.8])).pin_memory(),
                        "choose": torch.stack([torch.sort(torch.randperm(h * w, generator=g)[:500])[0] for _ in range(b)]).view(b, 1, 500).pin_memory(),
                        "obj": torch.randint(0, 21, (b,), generator=g).pin_memory()})
        return out
    batches = [host_batch(s) for s in range(5)]
    want = [pipe.estimate_buckets([{k: v.cuda() for k, v in bk.items()} for bk in hb]).cpu().clone() for hb in batches]
    se = StreamingEstimator(pipe, shapes)
    tickets = [se.submit(hb) for hb in batches[:2]]
    got = []
    for i in range(len(batches)):
        got.append(se.result(tickets[i]).clone())
        if i + 2 < len(batches):
            tickets.append(se.submit(batches[i + 2]))
    for g_, w_ in zip(got, want):
        assert torch.allclose(g_, w_, atol=1e-12, rtol=0), float((g_ - w_).abs().max())


def test_ragged_chunks_and_empty_inputs():
    """B not a multiple of the chunk (last chunk ragged), one-crop buckets, empty buckets / no detections at all."""
    import torch
    from densefusion_b200.pipeline import PoseEstimator
    est, ref, _, _ = build_nets(500, 21, seed=0)
    crops = [synth.synth_crop(20 + i, 500, 500, 21, (80, 80), obj=(3 * i) % 21) for i in range(5)]
    b = {k: torch.cat([c[k] for c in crops], 0).cuda() for k in ("img", "points", "choose", "idx")}
    whole = PoseEstimator(est, ref, iterations=2, precision="hybrid", chunk_crops=128)
    ragged = PoseEstimator(est, ref, iterations=2, precision="hybrid", chunk_crops=2)
    want = whole.estimate(b["img"], b["points"], b["choose"], b["idx"]).clone()
    got = ragged.estimate(b["img"], b["points"], b["choose"], b["idx"])
    assert torch.allclose(got, want, atol=1e-9, rtol=0)          # chunking only changes which rows share a launch
    single = ragged.estimate(b["img"][3:4], b["points"][3:4], b["choose"][3:4], b["idx"][3:4])
    # a lone crop takes the exact-fp32 kernel for the small-weight GEMMs that fall below 256 rows (ops.tc_eligible): same pose within parity
    assert float((single[0] - want[3]).abs().max()) < 1e-4 * float(want[3].abs().max())
    empty = {"img": b["img"][:0], "cloud": b["points"][:0], "choose": b["choose"][:0], "obj": b["idx"][:0].view(-1)}
    full = {"img": b["img"], "cloud": b["points"], "choose": b["choose"], "obj": b["idx"].view(-1)}
    assert whole.estimate_buckets([empty]).shape == (0, 7) and whole.estimate_buckets([]).shape == (0, 7)
    assert torch.allclose(whole.estimate_buckets([empty, full]), want, atol=1e-9, rtol=0)
    with pytest.raises(ValueError):
        whole.head_and_refine(b["points"][:, :400], torch.zeros(5 * 400, 32, device="cuda"), b["idx"].view(-1))
