"""Balanced schedule of the long-K convolutions (csrc/gemm_tc.cu, QSched; C-ABI df_conv_tc_schedule -- host-side arithmetic, no GPU):
every (tile, accumulation run) unit is given to exactly one cluster, ranges are contiguous and non-empty, a tile is shared by at
most two neighbouring clusters, and the busiest cluster has at least 10% less to do than under the round-robin tile walk."""
import ctypes
import itertools

import pytest

from densefusion_b200 import _C

WHOLE = 0x7FFF


def schedule(B, H, W, Cin, Cout, dil, clusters=74):
    out = (ctypes.c_int * (5 + 4 * clusters))()
    rc = _C.lib.df_conv_tc_schedule(B, H, W, Cin, Cout, dil, clusters, ctypes.cast(out, ctypes.c_void_p))
    return rc, list(out)


GEOMETRIES = [g for g in itertools.product((37, 64, 96, 128), (10, 15, 20), (256, 512), (256, 512), (1, 2, 4))]


@pytest.mark.parametrize("clusters", [74, 66])
def test_schedule_partitions_the_run_units(clusters):
    taken = 0
    for (B, HW, Cin, Cout, dil) in GEOMETRIES:
        rc, o = schedule(B, HW, HW, Cin, Cout, dil, clusters)
        assert rc in (0, 1)
        if rc != 1:
            continue
        taken += 1
        rr, sp, tiles, kbc, nkb_full = o[:5]
        assert sp * 100 <= rr * 90
        cl = min(tiles, clusters)
        rng = [tuple(o[5 + 4 * c: 9 + 4 * c]) for c in range(cl)]
        assert rng[0][:2] == (0, 0) and rng[-1][2:] == (tiles - 1, WHOLE)
        owners = {}
        for c, (t0, r0, t1, r1) in enumerate(rng):
            assert 0 <= t0 <= t1 < tiles and r0 >= 0 and r1 > 0
            assert not (t0 == t1 and r0 > 0 and r1 != WHOLE), "a tile cut twice"
            if t0 == t1 and r1 != WHOLE:
                assert r1 > r0
            for t in range(t0, t1 + 1):
                owners.setdefault(t, []).append(c)
            if c + 1 < cl:                                   # the next range starts exactly where this one stops
                nt0, nr0 = rng[c + 1][:2]
                assert (nt0, nr0) == ((t1 + 1, 0) if r1 == WHOLE else (t1, r1))
            max_runs = (nkb_full + kbc - 1) // kbc
            assert r0 < max_runs and (r1 == WHOLE or r1 < max_runs)
        assert sorted(owners) == list(range(tiles))
        for t, cs in owners.items():
            assert len(cs) <= 2 and (len(cs) == 1 or cs[1] == cs[0] + 1)
    assert taken > 20                                        # the schedule is actually used on these shapes


def test_schedule_rejects_bad_arguments_and_short_k():
    out = (ctypes.c_int * (5 + 4 * 74))()
    p = ctypes.cast(out, ctypes.c_void_p)
    assert _C.lib.df_conv_tc_schedule(0, 15, 15, 512, 512, 1, 74, p) < 0
    assert _C.lib.df_conv_tc_schedule(96, 15, 15, 500, 512, 1, 74, p) < 0
    assert _C.lib.df_conv_tc_schedule(96, 15, 15, 512, 512, 1, 200, p) < 0
    assert _C.lib.df_conv_tc_schedule(96, 15, 15, 128, 256, 1, 74, p) == 0      # one accumulation run per tile: nothing to cut
    assert _C.lib.df_conv_tc_schedule(8, 15, 15, 512, 512, 1, 74, p) == 0       # fewer tiles than clusters


def gemm_plan(M, N, K, groups=1, pooled=False, rows_per_crop=0, clusters=74):
    out = (ctypes.c_int * 6)()
    rc = _C.lib.df_gemm_tc_plan(M, N, K, groups, 1 if pooled else 0, rows_per_crop, clusters, ctypes.cast(out, ctypes.c_void_p))
    return rc, dict(zip(("width", "a_in_smem", "accumulators", "tiles", "rounds", "runs"), out))


def test_gemm_tile_plan_of_the_bench_shapes():
    """The launcher's tile width / operand placement for the head's GEMMs at the bench chunk (256 crops x 500 points), pinned as DESIGN.md
    section 4 states them (df_gemm_tc_plan: host-side arithmetic, the same functions the launch calls): tower-1 on 192-wide tiles with A in
    TMEM and two accumulators (1920 = 10 x 192; 256-wide tiles would need a half-masked tail tile -- DF_TC_TAIL256, measured slower),
    everything whose N is a multiple of 256 on 256-wide tiles with the A planes in shared memory and two accumulators."""
    rows = 256 * 500
    rc, t1 = gemm_plan(rows, 1920, 384)
    assert rc == 0 and t1 == {"width": 192, "a_in_smem": 0, "accumulators": 2, "tiles": 500 * 10, "rounds": 68, "runs": 1}
    for (N, K, groups) in ((256, 640, 3), (512, 384, 1), (512, 256, 1)):             # tower-2 (grouped), conv5 of the refiner / the head
        rc, p = gemm_plan(rows, N, K, groups)
        assert rc == 0 and (p["width"], p["a_in_smem"], p["accumulators"], p["runs"]) == (256, 1, 2, 1), (N, K, groups, p)
        assert p["tiles"] == 500 * (N // 256) * groups
    rc, c6 = gemm_plan(rows, 1024, 512, pooled=True, rows_per_crop=500)               # conv6 with the pooled epilogue: crop-aligned tiles
    assert rc == 0 and (c6["width"], c6["a_in_smem"], c6["accumulators"]) == (256, 1, 2) and c6["tiles"] == 256 * 2 * 4
    rc, up2 = gemm_plan(102400, 576, 256)                                             # decoder stage 2 at the low resolution: 3 x 192
    assert rc == 0 and (up2["width"], up2["a_in_smem"], up2["tiles"]) == (192, 0, 400 * 3)
    rc, t3 = gemm_plan(rows, 128, 256, 3)                                             # tower-3: one 128-wide tile per group
    assert rc == 0 and (t3["width"], t3["a_in_smem"]) == (128, 0)


def test_gemm_tile_plan_rejects_what_the_launcher_rejects():
    assert gemm_plan(1000, 1920, 100)[0] < 0                                          # K not a multiple of the 32-wide k-block
    assert gemm_plan(1000, 96, 128, groups=3)[0] < 0                                  # grouped layers: N % 128
    assert gemm_plan(1000, 1024, 512, pooled=True, rows_per_crop=300)[0] < 0          # pooled rows must be whole crops
