"""Colour encoder on the tensor cores: df_conv_tc against float64 convolutions, the NHWC helper
kernels against torch, and the whole encoder against the oracle's restatement of lib/pspnet.py + lib/extractors.py
(which tests/test_oracle_golden.py pins to the reference's own embeddings).  Bounds (max-abs / max-abs): whole encoder 3xtf32 <= 1e-4, tf32 <= 2e-2 (25 layers of single-pass TF32); one convolution
3xtf32 <= 2e-5 (the tensor core truncates while accumulating, see TcParams::k_chunks in gemm_tc.cu), tf32 <= 3e-3."""
import pytest
import torch
import torch.nn.functional as F

from densefusion_b200 import synth
from oracle import df_oracle as O
from densefusion_b200 import ops
from util import build_nets, rel

pytestmark = pytest.mark.gpu


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("B,H,W,Cin,Cout,taps,dil,extras", [
    (5, 10, 10, 256, 512, 9, 1, "relu"),            # 100-pixel maps, patches span crops
    (3, 10, 10, 512, 512, 9, 4, "residual"),        # layer4.1: dilation 4 reaches far outside a 10x10 map
    (2, 15, 15, 128, 256, 9, 2, "relu"),            # odd size
    (2, 20, 12, 64, 64, 9, 1, "residual"),          # non-square, 64 output channels
    (2, 40, 40, 256, 64, 9, 1, "prelu"),            # up_2
    (1, 80, 80, 64, 64, 9, 1, "prelu"),             # up_3
    (4, 10, 10, 256, 512, 1, 1, "none"),            # 1x1 projection
    (3, 20, 20, 1024, 256, 9, 1, "prelu"),          # up_1 (K = 9216)
    (5, 20, 20, 512, 512, 9, 4, "residual"),        # layer4.1 on a 160 px crop: border patches skip taps, odd patch count
    (7, 15, 15, 512, 512, 9, 4, "relu"),            # ... 120 px crop (3x3 patches)
    (33, 10, 10, 256, 256, 9, 2, "relu"),           # ... layer3.1, two crop groups per patch position
])
def test_conv_tc_vs_float64(B, H, W, Cin, Cout, taps, dil, extras):
    from densefusion_b200.encoder import PackedEncoder, _pack_conv
    g = torch.Generator().manual_seed(B * 1000 + H + Cin + Cout)
    k = 3 if taps == 9 else 1
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    bias = torch.randn(Cout, generator=g) if extras == "prelu" else None
    res = torch.randn(B, Cout, H, W, generator=g) if extras == "residual" else None
    slope = torch.tensor([0.25])
    want = F.conv2d(x.double(), w.double(), None if bias is None else bias.double(), padding=dil if k == 3 else 0, dilation=dil)
    if res is not None:
        want = want + res.double()
    if extras in ("relu", "residual"):
        want = torch.relu(want)
    elif extras == "prelu":
        want = torch.where(want > 0, want, 0.25 * want)
    for precision, tol in (("3xtf32", 2e-5), ("hybrid", 2e-5), ("hybrid16", 2e-5), ("hybrid16s", 2e-5), ("tf32", 3e-3)):
        wide = torch.full((B, H, W, Cout + 64), 7.0, device="cuda")          # output is a channel slice of a wider buffer
        out = wide[..., 32:32 + Cout]
        PackedEncoder._conv(_nhwc(x).cuda(), _pack_conv(w.cuda()), out, taps=taps, dil=dil,
                            bias=None if bias is None else bias.cuda(), residual=None if res is None else _nhwc(res).cuda(),
                            prelu=slope.cuda() if extras == "prelu" else None,
                            act={"relu": 1, "residual": 1, "prelu": 2, "none": 0}[extras],
                            mode=ops.PRECISIONS[precision])
        torch.cuda.synchronize()
        err = rel(out.permute(0, 3, 1, 2), want)
        print(f"conv {taps}tap dil{dil} {Cin}->{Cout} {H}x{W}x{B} {precision}: {err:.3e}")
        assert err < tol
        assert float(wide[..., :32].min()) == 7.0 and float(wide[..., 32 + Cout:].max()) == 7.0    # neighbours untouched


@pytest.mark.parametrize("B,H,W,Cin,Cout,dil,extras", [
    (37, 15, 15, 512, 512, 2, "residual"),          # 70 tiles of <= 144 k-blocks on 74 CTA pairs... see the assertion on the schedule
    (96, 15, 15, 512, 512, 4, "relu"),              # layer4.1 of the bench step's 120 px bucket: border tiles skip taps (ragged runs)
    (40, 20, 20, 512, 512, 1, "residual"),          # 126 tiles
    (64, 20, 20, 256, 256, 2, "relu"),              # two runs per tile
])
def test_conv_tc_balanced_schedule_vs_float64(B, H, W, Cin, Cout, dil, extras):
    """Launches large enough to take the balanced (tile, run) schedule of gemm_tc.cu (QSched): tiles shared by two clusters are
    handed over through the per-warp flags; the result must match a float64 convolution like any other launch, be the same from
    run to run, and leave the neighbouring channels of a wider buffer alone."""
    import ctypes
    from densefusion_b200 import _C
    from densefusion_b200.encoder import PackedEncoder, _pack_conv
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    info = (ctypes.c_int * (5 + 4 * 80))()
    planned = _C.lib.df_conv_tc_schedule(B, H, W, Cin, Cout, dil, min(sms // 2, 80), ctypes.cast(info, ctypes.c_void_p))
    assert planned == 1, "the shape no longer takes the balanced schedule: pick another one"
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H + Cin + Cout)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (Cin * 9) ** 0.5
    bias = torch.randn(Cout, device="cuda", generator=g)
    res = torch.randn(B, H, W, Cout, device="cuda", generator=g) if extras == "residual" else None
    want = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), bias.double(), padding=dil, dilation=dil).permute(0, 2, 3, 1)
    if res is not None:
        want = want + res.double()
    want = torch.relu(want)
    outs = []
    for rep in range(3):
        wide = torch.full((B, H, W, Cout + 64), 7.0, device="cuda")
        out = wide[..., 32:32 + Cout]
        PackedEncoder._conv(x, _pack_conv(w), out, taps=9, dil=dil, bias=bias, residual=res, act=1, mode=ops.PRECISIONS["hybrid16s"])
        torch.cuda.synchronize()
        assert float(wide[..., :32].min()) == 7.0 and float(wide[..., 32 + Cout:].max()) == 7.0
        outs.append(out.clone())
    err = float((outs[0].double() - want).abs().max() / want.abs().max())
    print(f"balanced schedule {B}x{H}x{W} {Cin}->{Cout} dil {dil}: {err:.3e} (busiest cluster {info[1]} instead of {info[0]} k-blocks)")
    assert err < 2e-5
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_encoder_helper_kernels_vs_torch():
    from densefusion_b200._C import check, lib, ptr, stream
    g = torch.Generator().manual_seed(2)
    B, H, W, C = 3, 40, 24, 64
    x = torch.randn(B, C, H, W, generator=g)
    xn = _nhwc(x).cuda()
    # max pool 3x3/2
    out = torch.empty(B, H // 2, W // 2, C, device="cuda")
    check(lib.df_enc_maxpool(ptr(xn), ptr(out), B, H, W, C, stream()), "maxpool")
    assert torch.equal(out.permute(0, 3, 1, 2).cpu(), F.max_pool2d(x, 3, 2, 1))
    # im2col of the 3x3/2 patches == unfold (tap-major, channels fastest)
    A = torch.empty(B * (H // 2) * (W // 2), 9 * C, device="cuda")
    check(lib.df_enc_im2col_s2(ptr(xn), ptr(A), B, H, W, C, stream()), "im2col_s2")
    unf = F.unfold(x, 3, padding=1, stride=2).view(B, C, 9, -1).permute(0, 3, 2, 1).reshape(-1, 9 * C)
    assert torch.equal(A.cpu(), unf)
    # conv1 im2col
    img = torch.randn(2, 3, 40, 56, generator=g)
    A1 = torch.empty(2 * 20 * 28, 160, device="cuda")
    check(lib.df_enc_im2col_conv1(ptr(img.cuda()), ptr(A1), 2, 40, 56, 160, stream()), "im2col_conv1")
    unf1 = F.unfold(img, 7, padding=3, stride=2).permute(0, 2, 1).reshape(-1, 147)
    assert torch.equal(A1[:, :147].cpu(), unf1) and float(A1[:, 147:].abs().max()) == 0.0
    # adaptive average pooling from a channel slice of a wider buffer
    wide = torch.randn(B, 10, 15, 96, generator=g).cuda()
    for S in (1, 2, 3, 6):
        o = torch.empty(B, S, S, 64, device="cuda")
        check(lib.df_enc_adaptive_avgpool(ptr(wide[..., 32:]), 96, ptr(o), B, 10, 15, 64, S, stream()), "avgpool")
        want = F.adaptive_avg_pool2d(wide[..., 32:].permute(0, 3, 1, 2).cpu(), (S, S))
        assert rel(o.permute(0, 3, 1, 2), want) < 1e-6
    # folded pyramid: the four pools in one pass (stage-major rows), and the sum of their bilinear resizes
    pooled = torch.empty(50 * B, 64, device="cuda")
    check(lib.df_enc_pyramid_pool(ptr(wide[..., 32:]), 96, ptr(pooled), B, 10, 15, 64, stream()), "pyramid_pool")
    src = wide[..., 32:].permute(0, 3, 1, 2).cpu()
    row, want_sum = 0, torch.zeros(B, 64, 10, 15)
    for S in (1, 2, 3, 6):
        want = F.adaptive_avg_pool2d(src, (S, S))
        got = pooled[row:row + B * S * S].view(B, S, S, 64).permute(0, 3, 1, 2)
        assert rel(got, want) < 1e-6
        want_sum += F.interpolate(want, size=(10, 15), mode="bilinear", align_corners=False)
        row += B * S * S
    summed = torch.zeros(B, 10, 15, 96, device="cuda")
    check(lib.df_enc_pyramid_sum(ptr(pooled), ptr(summed[..., 16:]), 96, B, 10, 15, 64, stream()), "pyramid_sum")
    assert rel(summed[..., 16:80].permute(0, 3, 1, 2), want_sum) < 2e-6
    assert float(summed[..., :16].abs().max()) == 0.0 and float(summed[..., 80:].abs().max()) == 0.0
    # decoder stage at the low resolution: tap products (here in fp64 on the host) + df_enc_upconv_finish
    #   == PReLU(conv3x3(resize x2 (align_corners)) + bias)
    xl, wt = torch.randn(2, 32, 6, 9, generator=g), torch.randn(16, 32, 3, 3, generator=g) * 0.1
    bs, slope = torch.randn(16, generator=g), torch.tensor([0.25])
    want = F.prelu(F.conv2d(F.interpolate(xl, scale_factor=2, mode="bilinear", align_corners=True), wt, bs, padding=1), slope)
    z = (_nhwc(xl).reshape(-1, 32).double() @ wt.permute(2, 3, 0, 1).reshape(144, 32).double().T).float().view(2, 6, 9, 144).cuda()
    o = torch.empty(2, 12, 18, 16, device="cuda")
    bs_d, slope_d = bs.cuda(), slope.cuda()                  # (named: a temporary would be freed before the launch)
    check(lib.df_enc_upconv_finish(ptr(z), 144, ptr(bs_d), ptr(slope_d), ptr(o), 16, 2, 6, 9, 16, stream()), "upconv_finish")
    assert rel(o.permute(0, 3, 1, 2), want) < 5e-6
    # bilinear resize into a channel slice, both alignment modes
    small = torch.randn(B, 6, 6, 64, generator=g).cuda()
    for (ho, wo, align) in ((10, 15, False), (12, 12, True)):
        dst = torch.zeros(B, ho, wo, 160, device="cuda")
        check(lib.df_enc_upsample(ptr(small), 64, ptr(dst[..., 64:]), 160, B, 6, 6, ho, wo, 64, 1 if align else 0, stream()), "up")
        want = F.interpolate(small.permute(0, 3, 1, 2).cpu(), size=(ho, wo), mode="bilinear", align_corners=align)
        assert rel(dst[..., 64:128].permute(0, 3, 1, 2), want) < 2e-6 and float(dst[..., :64].abs().max()) == 0.0
    # log-softmax over 32 channels
    z = torch.randn(1000, 32, generator=g).cuda()
    want = torch.log_softmax(z.cpu().double(), 1)
    check(lib.df_enc_log_softmax32(ptr(z), 1000, stream()), "lsm")
    assert rel(z, want) < 1e-6


@pytest.mark.parametrize("precision,tol", [("3xtf32", 1e-4), ("hybrid", 1e-4), ("hybrid16", 1e-4), ("hybrid16s", 1e-4), ("tf32", 2e-2)])
@pytest.mark.parametrize("b,hw", [(3, (80, 80)), (2, (120, 160)), (1, (160, 160))])
def test_encoder_vs_oracle(precision, tol, b, hw):
    from densefusion_b200.encoder import PackedEncoder
    est, _, est_sd, _ = build_nets(500, 21, seed=6)
    g = torch.Generator().manual_seed(hw[0] + b)
    img = torch.randn(b, 3, hw[0], hw[1], generator=g)
    enc = PackedEncoder(est.cnn)
    feat = enc.forward(img.cuda(), precision)                      # (b,H,W,32)
    torch.cuda.synchronize()
    with torch.no_grad():
        want = O.psp_encoder(est_sd, img)
    err = rel(feat.permute(0, 3, 1, 2), want)
    print(f"encoder {precision} {b}x{hw}: rel err {err:.3e}")
    assert err < tol


@pytest.mark.parametrize("b,hw", [(3, (80, 80)), (2, (120, 160))])
def test_encoder_sparse_tail_equals_dense_gather(b, hw):
    """forward_points (last decoder stage on gathered 3x3 patches of the chosen pixels only) == dense encoder + gather,
    including pixels on the image border (zero padding of the 3x3 convolution) -- and both match the oracle."""
    from densefusion_b200 import ops
    from densefusion_b200.encoder import PackedEncoder
    est, _, est_sd, _ = build_nets(500, 21, seed=6)
    g = torch.Generator().manual_seed(hw[1] + b)
    img = torch.randn(b, 3, hw[0], hw[1], generator=g)
    H, W = hw
    choose = torch.stack([torch.sort(torch.randperm(H * W, generator=g)[:500])[0] for _ in range(b)])
    choose[:, 0], choose[:, 1], choose[:, 2], choose[:, 3] = 0, W - 1, (H - 1) * W, H * W - 1      # the four corners
    enc = PackedEncoder(est.cnn)
    dense = enc.forward(img.cuda(), "3xtf32")
    pm_dense, _ = ops.gather_embedding(dense.permute(0, 3, 1, 2), choose.cuda(), want_cm=False)
    pm = torch.empty(b * 500, 32, device="cuda")
    enc.forward_points(img.cuda(), choose.cuda(), pm, "3xtf32")
    torch.cuda.synchronize()
    err = rel(pm, pm_dense)
    with torch.no_grad():
        want = O.gather_embedding(O.psp_encoder(est_sd, img), choose.view(b, 1, 500))          # (b,32,N)
    err_o = rel(pm.view(b, 500, 32).permute(0, 2, 1), want)
    print(f"sparse tail vs dense gather {err:.3e}; vs oracle {err_o:.3e}")
    assert err < 1e-5 and err_o < 1e-4


@pytest.mark.parametrize("precision", ["3xtf32", "hybrid", "hybrid16", "hybrid16s"])
def test_pipeline_with_tensor_core_encoder_vs_oracle(precision):
    """estimate + 2 refine iterations with encoder AND head on the tensor cores (fp32-parity modes) against the oracle's
    eval loop."""
    import numpy as np
    from densefusion_b200.pipeline import PoseEstimator
    est, ref, est_sd, ref_sd = build_nets(500, 21, seed=0)
    d = synth.synth_crop(7, 500, 500, 21, (80, 80), obj=12)
    dc = {k: v.cuda() for k, v in d.items()}
    pipe = PoseEstimator(est, ref, iterations=2, precision=precision)
    assert pipe.encoder == "tc"
    pose = pipe.estimate(dc["img"], dc["points"], dc["choose"], dc["idx"]).cpu().numpy()[0]
    want = O.estimate_and_refine(est_sd, ref_sd, d["img"], d["points"], d["choose"], d["idx"], 21, 2)
    err = float(np.max(np.abs(pose - want)) / np.max(np.abs(want)))
    print(f"pose err with tensor-core encoder ({precision}): {err:.3e}")
    assert err < 1e-4


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,dil,bias", [(3, 10, 10, 256, 512, 3, 4, False), (2, 20, 20, 128, 128, 3, 1, False),
                                                       (2, 40, 40, 256, 64, 3, 1, True), (4, 10, 10, 2560, 1024, 1, 1, True),
                                                       (2, 80, 80, 64, 32, 1, 1, True), (3, 15, 15, 256, 256, 3, 2, False)])
def test_training_convolution_forward_and_gradients_vs_float64(B, H, W, Cin, Cout, k, dil, bias):
    """lib.conv_tc.ConvTCFn (tensor-core forward + data gradient, library weight gradient) against float64 autograd."""
    import torch.nn as nn
    from densefusion_b200.lib import conv_tc
    g = torch.Generator().manual_seed(B + H + Cin + Cout)
    m = nn.Conv2d(Cin, Cout, k, padding=dil * (k // 2), dilation=dil, bias=bias)
    with torch.no_grad():
        m.weight.copy_(torch.randn(m.weight.shape, generator=g) / (Cin * k * k) ** 0.5)
        if bias:
            m.bias.copy_(torch.randn(Cout, generator=g))
    x = torch.randn(B, Cin, H, W, generator=g)
    gy = torch.randn(B, Cout, H, W, generator=g)
    m64 = nn.Conv2d(Cin, Cout, k, padding=dil * (k // 2), dilation=dil, bias=bias).double()
    m64.load_state_dict({kk: v.double() for kk, v in m.state_dict().items()})
    x64 = x.double().requires_grad_(True)
    y64 = m64(x64)
    y64.backward(gy.double())
    m = m.cuda()
    torch.backends.cudnn.allow_tf32 = False
    xc = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    assert conv_tc.eligible(m, xc)
    y = conv_tc.conv2d(m, xc)
    y.backward(gy.cuda())
    errs = (rel(y, y64), rel(xc.grad, x64.grad), rel(m.weight.grad, m64.weight.grad))
    print(f"conv_tc {k}x{k} dil{dil} {Cin}->{Cout} {H}x{W}x{B}: fwd {errs[0]:.2e} dx {errs[1]:.2e} dW {errs[2]:.2e}")
    assert errs[0] < 2e-5 and errs[1] < 2e-5 and errs[2] < 2e-2          # dW: cuDNN's fp32 weight-gradient algorithms
    if bias:
        assert rel(m.bias.grad, m64.bias.grad) < 1e-5


@pytest.mark.parametrize("hin,win,hout,wout,align", [(10, 10, 20, 20, True), (1, 1, 15, 20, False), (2, 2, 20, 20, False),
                                                     (3, 3, 10, 15, False), (6, 6, 20, 20, False), (15, 20, 30, 40, True)])
def test_upsample_backward_kernel_vs_torch_autograd(hin, win, hout, wout, align):
    from densefusion_b200._C import check, lib, ptr, stream
    g = torch.Generator().manual_seed(hin + hout)
    B, C = 3, 64
    x = torch.randn(B, C, hin, win, generator=g, dtype=torch.float64, requires_grad=True)
    gy = torch.randn(B, C, hout, wout, generator=g)
    F.interpolate(x, size=(hout, wout), mode="bilinear", align_corners=align).backward(gy.double())
    gy_n = gy.cuda().contiguous(memory_format=torch.channels_last)
    gi = torch.empty(B, C, hin, win, device="cuda", memory_format=torch.channels_last)
    check(lib.df_enc_upsample_backward(ptr(gy_n), C, ptr(gi), C, B, hin, win, hout, wout, C, 1 if align else 0, stream()), "up_bwd")
    assert rel(gi, x.grad) < 2e-6


@pytest.mark.parametrize("B,H,W,Cin,Cout,k", [(2, 40, 56, 3, 64, 7), (3, 20, 20, 64, 128, 3), (2, 30, 22, 64, 128, 3),
                                              (3, 20, 20, 64, 128, 1), (1, 15, 11, 64, 128, 1)])
def test_stride2_convolutions_forward_and_gradients_vs_float64(B, H, W, Cin, Cout, k):
    """lib.conv_tc.ConvS2Fn (im2col + tensor-core GEMM, col2im, tensor-core weight gradient) against float64 autograd."""
    import torch.nn as nn
    from densefusion_b200.lib import conv_tc
    g = torch.Generator().manual_seed(B + H + Cin + k)
    m = nn.Conv2d(Cin, Cout, k, stride=2, padding=k // 2, bias=False)
    with torch.no_grad():
        m.weight.copy_(torch.randn(m.weight.shape, generator=g) / (Cin * k * k) ** 0.5)
    x = torch.randn(B, Cin, H, W, generator=g)
    m64 = nn.Conv2d(Cin, Cout, k, stride=2, padding=k // 2, bias=False).double()
    m64.weight.data.copy_(m.weight.data.double())
    x64 = x.double().requires_grad_(k != 7)
    y64 = m64(x64)
    dy = torch.randn(y64.shape, generator=g)
    (y64 * dy.double()).sum().backward()
    m = m.cuda()
    xc = x.cuda().requires_grad_(k != 7)
    old = conv_tc.STRIDE2_TC
    conv_tc.STRIDE2_TC = True
    try:
        assert conv_tc.eligible_s2(m, xc)
        y = conv_tc.conv2d(m, xc)
        (y * dy.cuda()).sum().backward()
    finally:
        conv_tc.STRIDE2_TC = old
    torch.cuda.synchronize()
    e = [rel(y, y64), rel(m.weight.grad, m64.weight.grad)] + ([rel(xc.grad, x64.grad)] if k != 7 else [])
    print(f"conv_s2 {k}x{k} {Cin}->{Cout} {H}x{W}x{B}: " + " ".join(f"{v:.2e}" for v in e))
    assert max(e) < 2e-5


@pytest.mark.parametrize("B,C,h,w", [(2, 64, 6, 9), (3, 32, 7, 7), (1, 256, 10, 10), (2, 64, 2, 3)])
def test_upconv_finish_staged_kernel_vs_conv_of_resized_map(B, C, h, w):
    """Decoder stage at the low resolution on the shared-memory kernel (C % 32 == 0): tap products + df_enc_upconv_finish ==
    PReLU(conv3x3(resize x2, align_corners) + bias), including patches on every border of the map and a channel-slice output."""
    from densefusion_b200._C import check, lib, ptr, stream
    g = torch.Generator().manual_seed(B * 1000 + C + h * w)
    cin = 16
    xl, wt = torch.randn(B, cin, h, w, generator=g), torch.randn(C, cin, 3, 3, generator=g) * 0.1
    bs, slope = torch.randn(C, generator=g), torch.tensor([0.25])
    want = F.prelu(F.conv2d(F.interpolate(xl.double(), scale_factor=2, mode="bilinear", align_corners=True), wt.double(), bs.double(), padding=1),
                   slope.double())
    z = (_nhwc(xl).reshape(-1, cin).double() @ wt.permute(2, 3, 0, 1).reshape(9 * C, cin).double().T).float().view(B, h, w, 9 * C).cuda()
    o = torch.full((B, 2 * h, 2 * w, C + 32), 7.0, device="cuda")
    bs_d, slope_d = bs.cuda(), slope.cuda()
    check(lib.df_enc_upconv_finish(ptr(z), 9 * C, ptr(bs_d), ptr(slope_d), ptr(o), C + 32, B, h, w, C, stream()), "upconv_finish")
    assert rel(o[..., :C].permute(0, 3, 1, 2), want) < 5e-6
    assert float((o[..., C:] - 7.0).abs().max()) == 0.0


@pytest.mark.parametrize("B,H,W", [(2, 80, 80), (3, 40, 56), (1, 24, 16)])
def test_conv1_gathered_by_the_kernel_vs_conv2d(B, H, W):
    """df_enc_conv1_tc (7x7 / stride 2 / pad 3 with the patches gathered by the GEMM kernel's stagers, hybrid16s) == relu(conv2d) in float64,
    borders and the masked tail of the last 256-row tile included; and == the im2col + GEMM path to rounding."""
    from densefusion_b200._C import check, lib, ptr, stream
    g = torch.Generator().manual_seed(B + H + W)
    img = torch.randn(B, 3, H, W, generator=g)
    w = torch.randn(64, 3, 7, 7, generator=g) * (2.0 / (49 * 64)) ** 0.5
    want = torch.relu(F.conv2d(img.double(), w.double(), stride=2, padding=3))
    w1 = torch.zeros(64, 160)
    w1[:, :147] = w.reshape(64, 147)
    sw = ops.SplitWeight(w1.cuda())
    planes, scale = sw.planes16s()
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.full((B, Ho, Wo, 96), 7.0, device="cuda")
    imgc = img.cuda()
    check(lib.df_enc_conv1_tc(ptr(imgc), B, H, W, ptr(planes), ptr(scale), ptr(y), 96, 64, 1, stream()), "df_enc_conv1_tc")
    assert rel(y[..., :64].permute(0, 3, 1, 2), want) < 5e-6
    assert float((y[..., 64:] - 7.0).abs().max()) == 0.0
    a0 = torch.empty(B * Ho * Wo, 160, device="cuda")
    check(lib.df_enc_im2col_conv1(ptr(imgc), ptr(a0), B, H, W, 160, stream()), "df_enc_im2col_conv1")
    y2 = torch.empty(B * Ho * Wo, 64, device="cuda")
    ops.gemm(a0, sw, None, y2, M=B * Ho * Wo, N=64, K=160, lda=160, ldw=160, ldc=64, relu=True, precision="hybrid16s")
    assert rel(y[..., :64].reshape(-1, 64), y2) < 3e-6
