"""CPU-only: the C-ABI library loads and exports every symbol include/densefusion_b200.h declares, and the
ctypes table covers exactly that set (no compute calls)."""
import ctypes
import os
import re

from util import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "densefusion_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|long long)\s+(df_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from densefusion_b200 import _C
    names = header_functions()
    assert len(names) >= 15
    lib = ctypes.CDLL(_C.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_C.SIGNATURES) == names


def test_version_and_features():
    from densefusion_b200 import _C
    assert _C.lib.df_abi_version() == 1
    assert _C.lib.df_features() & 1          # tcgen05 path compiled in


def test_argument_errors_do_not_need_a_gpu():
    from densefusion_b200 import _C
    assert _C.lib.df_knn(None, None, None, 1, 3, 10, 10, 1, None) == -1
    assert _C.lib.df_pose_compose(None, None, None, 1, None) == -1


def test_state_dict_keys_match_reference_layout():
    """Keys/shapes the reference checkpoints use; spot list + counts."""
    from densefusion_b200.lib.network import PoseNet, PoseRefineNet
    sd = PoseNet(500, 13).state_dict()
    assert len(sd) == 77
    assert tuple(sd["cnn.model.module.feats.conv1.weight"].shape) == (64, 3, 7, 7)
    assert tuple(sd["cnn.model.module.psp.stages.3.1.weight"].shape) == (512, 512, 1, 1)
    assert tuple(sd["cnn.model.module.up_1.conv.2.weight"].shape) == (1,)
    assert tuple(sd["cnn.model.module.classifier.2.weight"].shape) == (21, 256)
    assert tuple(sd["feat.conv1.weight"].shape) == (64, 3, 1)
    assert tuple(sd["conv1_r.weight"].shape) == (640, 1408, 1)
    assert tuple(sd["conv4_r.weight"].shape) == (52, 128, 1)
    rd = PoseRefineNet(500, 13).state_dict()
    assert len(rd) == 24
    assert tuple(rd["feat.conv5.weight"].shape) == (512, 384, 1)
    assert tuple(rd["conv1_r.weight"].shape) == (512, 1024)
    assert tuple(rd["conv3_t.weight"].shape) == (39, 128)


def test_state_dict_keys_equal_reference_when_available():
    import sys
    if not os.path.isdir("/root/reference/lib"):
        import pytest
        pytest.skip("reference tree not present (GPU box)")
    import importlib
    import warnings
    warnings.filterwarnings("ignore")
    sys.path.insert(0, "/root/reference")
    try:
        saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "lib" or k.startswith("lib.")}
        ref_net = importlib.import_module("lib.network")
        ours = importlib.import_module("densefusion_b200.lib.network")
        for cls in ("PoseNet", "PoseRefineNet"):
            a = {k: tuple(v.shape) for k, v in getattr(ref_net, cls)(500, 21).state_dict().items()}
            b = {k: tuple(v.shape) for k, v in getattr(ours, cls)(500, 21).state_dict().items()}
            assert a == b
    finally:
        sys.path.remove("/root/reference")
        for k in list(sys.modules):
            if k == "lib" or k.startswith("lib."):
                del sys.modules[k]
        sys.modules.update(saved)


def test_wgrad_scratch_size_is_a_pure_host_function():
    """df_conv_wgrad_scratch_floats: geometry only (padded row length a multiple of 4, pixel axis cut into 32-aligned
    slices, three pre-shifted hi/lo planes of X for 3x3, the split-K partial outputs) -- callable without a GPU."""
    from densefusion_b200 import _C
    f = _C.lib.df_conv_wgrad_scratch_floats
    # 1x1: no padding, one plane pair of X; P = 4*10*10 = 400 -> 13 k-blocks, too short to split
    n = f(4, 10, 10, 512, 1024, 1, 1)
    assert n == 416 * (1024 + 2 * 512)
    # 3x3 dilation 4 on 10x10: padded rows of 20 (18 -> multiple of 4), 3 shifted hi/lo planes, split-K partials on top
    B, H, W, cin, cout, d = 6, 10, 10, 512, 512, 4
    n = f(B, H, W, cin, cout, 9, d)
    P = B * (H + 2 * d) * 20
    planes = cout + 2 * 3 * cin
    assert n >= ((P + 31) // 32 * 32) * planes and (n - 0) % 32 == 0
    # monotone in every extent
    assert f(B + 1, H, W, cin, cout, 9, d) > n and f(B, H, W, cin, 2 * cout, 9, d) > n
