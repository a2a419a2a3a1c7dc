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
