"""Input preparation on the device against the oracle's numpy restatement of
tools/eval_ycb.py:54-91 / :147-190.  Bit-exact for the box, the colour crop, and -- whenever the object has at most
num_points valid pixels (the reference's deterministic wrap-padding branch) -- `choose` and the cloud; with more pixels
the reference samples with numpy's RNG, so the test checks the properties of a valid sample instead."""
import numpy as np
import pytest
import torch

from oracle import df_oracle as O


def _frame(seed, H=480, W=640):
    rng = np.random.RandomState(seed)
    rgb = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
    depth = rng.randint(0, 20000, size=(H, W)).astype(np.float32)
    depth[rng.rand(H, W) < 0.2] = 0.0                                   # sensor holes
    label = np.zeros((H, W), dtype=np.int32)
    return rng, rgb, depth, label


def test_bbox_snapping_matches_oracle_on_random_and_edge_boxes():
    from densefusion_b200.cropper import get_bbox
    rng = np.random.RandomState(0)
    rois = [[0, 1, 10, 20, 51, 61], [0, 1, 0, 0, 639, 479], [0, 1, 600, 440, 639, 479], [0, 1, 5, 5, 46, 46],
            [0, 1, 100, 100, 181, 141], [0, 1, -5, -5, 30, 30]]
    for _ in range(300):
        x1, y1 = rng.randint(-10, 600), rng.randint(-10, 440)
        rois.append([0, 1, x1, y1, x1 + rng.randint(5, 400), y1 + rng.randint(5, 400)])
    for roi in rois:
        assert get_bbox(roi) == O.crop_bbox(roi), roi


@pytest.mark.gpu
def test_crops_match_reference_numpy_path():
    from densefusion_b200.cropper import CropBuilder
    N = 500
    rng, rgb, depth, label = _frame(1)
    # object 5: small blob (fewer than N valid pixels -> wrap padding); object 9: large (more than N -> sampling);
    # object 3: box clipped by the image border; object 7: not visible at all
    label[100:118, 200:222] = 5
    label[250:330, 300:420] = 9
    label[0:30, 0:45] = 3
    objects = [(0, 5, [0, 5, 195, 95, 230, 125]), (0, 9, [0, 9, 295, 245, 425, 335]), (0, 3, [0, 3, -3, -2, 50, 34]),
               (0, 7, [0, 7, 400, 50, 470, 120])]
    cb = CropBuilder(N)
    buckets = cb.build(torch.from_numpy(rgb)[None].cuda(), torch.from_numpy(depth)[None].cuda(),
                       torch.from_numpy(label)[None].cuda(), objects, seed=3)
    seen = 0
    for bk in buckets:
        for i, pos in enumerate(bk["order"]):
            frame, item, roi = objects[pos]
            cloud, choose, img, count = O.build_crop(rgb, depth, label, roi, item, N) if item != 7 else (None, None, None, 0)
            seen += 1
            assert int(bk["count"][i]) == count
            assert int(bk["obj"][i]) == item - 1
            if item == 7:
                assert float(bk["cloud"][i].abs().max()) == 0.0 and int(bk["choose"][i].max()) == 0
                continue
            assert torch.equal(bk["img"][i].cpu(), torch.from_numpy(img)), "colour crop must be bit-exact"
            got_choose = bk["choose"][i, 0].cpu().numpy()
            got_cloud = bk["cloud"][i].cpu().numpy()
            rmin, rmax, cmin, cmax = O.crop_bbox(roi)
            if count <= N:
                assert np.array_equal(got_choose, choose) and np.array_equal(got_cloud, cloud)
            else:
                valid = ((label == item) & (depth != 0))[rmin:rmax, cmin:cmax].flatten().nonzero()[0]
                assert len(np.unique(got_choose)) == N and np.all(np.diff(got_choose) > 0) and np.isin(got_choose, valid).all()
                rows, cols = got_choose // (cmax - cmin) + rmin, got_choose % (cmax - cmin) + cmin
                d = depth[rows, cols]
                pt2 = d / np.float32(10000.0)
                want = np.stack([(cols.astype(np.float32) - np.float32(312.9869)) * pt2 / np.float32(1066.778),
                                 (rows.astype(np.float32) - np.float32(241.3109)) * pt2 / np.float32(1067.487), pt2], 1)
                assert np.array_equal(got_cloud, want.astype(np.float32))
                # the sample is spread over the whole object, not a prefix
                assert got_choose[-1] > valid[int(0.9 * len(valid))]
    assert seen == len(objects)


@pytest.mark.gpu
def test_built_crops_feed_the_pose_pipeline():
    from densefusion_b200.cropper import CropBuilder
    from densefusion_b200.pipeline import PoseEstimator
    from util import build_nets
    rng, rgb, depth, label = _frame(2)
    label[100:170, 200:270] = 4
    label[300:400, 100:250] = 12
    objects = [(0, 4, [0, 4, 195, 95, 275, 175]), (0, 12, [0, 12, 95, 295, 255, 405])]
    buckets = CropBuilder(500).build(torch.from_numpy(rgb)[None].cuda(), torch.from_numpy(depth)[None].cuda(),
                                     torch.from_numpy(label)[None].cuda(), objects)
    est, ref, _, _ = build_nets(500, 21, seed=0)
    poses = PoseEstimator(est, ref, iterations=2, precision="hybrid").estimate_buckets(buckets)
    assert poses.shape == (2, 7) and torch.isfinite(poses).all()
    assert torch.allclose(poses[:, :4].norm(dim=1), torch.ones(2, dtype=torch.float64, device="cuda"), atol=1e-6)
