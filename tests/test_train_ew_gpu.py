"""The element-wise / pooling ops of the TRAINING encoder on own kernels (densefusion_b200/lib/ew.py, csrc/train_ew.cu) against the
torch ops the reference's module graph uses (lib/extractors.py:78-124, lib/pspnet.py:7-77): forward values and gradients."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CL = torch.channels_last


def _close(a, b, tol=1e-6):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape
    err = float((a - b).abs().max() / (b.abs().max() + 1e-30))
    assert err <= tol, err


torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _rand(*shape, seed=0, cl=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*shape, generator=g).cuda()
    return x.contiguous(memory_format=CL) if cl and len(shape) == 4 else x


@pytest.mark.parametrize("shape", [(3, 64, 40, 40), (2, 64, 41, 37), (1, 8, 5, 6)])
def test_maxpool_forward_backward_incl_ties(shape):
    from densefusion_b200.lib import ew
    x = _rand(*shape, seed=1)
    x = torch.round(x * 2) / 2                       # many exact ties inside a window: the first maximum must take the gradient
    x = x.contiguous(memory_format=CL).requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    y, yr = ew.MaxPoolFn.apply(x), F.max_pool2d(xr, 3, stride=2, padding=1)
    g = _rand(*yr.shape, seed=2)
    y.backward(g)
    yr.backward(g)
    _close(y, yr, 0.0)
    _close(x.grad, xr.grad, 0.0)


@pytest.mark.parametrize("b,h,w", [(3, 20, 20), (2, 15, 15), (2, 10, 13)])
def test_pyramid_pool_and_concat_vs_torch(b, h, w):
    from densefusion_b200.lib import ew
    c = 64
    f = _rand(b, c, h, w, seed=3).requires_grad_(True)
    fr = f.detach().clone().requires_grad_(True)
    ws = [_rand(c, c, 1, 1, seed=10 + i, cl=False) * 0.1 for i in range(4)]
    pooled = ew.PyramidPoolFn.apply(f)
    ys = [F.conv2d(p, wt) for p, wt in zip(pooled, ws)]
    cat = ew.PyramidCatFn.apply(f, *ys)
    pr = [F.adaptive_avg_pool2d(fr, (s, s)) for s in (1, 2, 3, 6)]
    for p, q in zip(pooled, pr):
        _close(p, q, 2e-6)
    catr = torch.cat([F.interpolate(F.conv2d(p, wt), size=(h, w), mode="bilinear", align_corners=False) for p, wt in zip(pr, ws)] + [fr], 1)
    _close(cat, catr, 3e-6)
    g = _rand(*catr.shape, seed=4)
    cat.backward(g)
    catr.backward(g)
    _close(f.grad, fr.grad, 2e-5)                    # (summation order of the overlapping bins and of the four branches)


def test_prelu_dropout_logsoftmax_vs_torch():
    from densefusion_b200.lib import ew
    x = _rand(2, 64, 24, 24, seed=5).requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    a = torch.tensor([0.25], device="cuda", requires_grad=True)
    ar = a.detach().clone().requires_grad_(True)
    y, yr = ew.PReLUFn.apply(x, a), F.prelu(xr, ar)
    g = _rand(*yr.shape, seed=6)
    y.backward(g)
    yr.backward(g)
    _close(y, yr, 0.0)
    _close(x.grad, xr.grad, 0.0)
    _close(a.grad, ar.grad, 1e-5)

    z = _rand(3, 32, 16, 16, seed=7).requires_grad_(True)
    zr = z.detach().clone().requires_grad_(True)
    y, yr = ew.LogSoftmax32Fn.apply(z), F.log_softmax(zr, dim=1)
    g = _rand(*yr.shape, seed=8)
    y.backward(g)
    yr.backward(g)
    _close(y, yr, 2e-6)
    _close(z.grad, zr.grad, 5e-6)

    # Dropout2d: whole (sample, channel) maps dropped, survivors scaled by 1 / (1 - p); backward uses the same mask; masks change from
    # call to call (the device-side counter) and the drop rate is about p
    d = torch.ones(64, 256, 4, 4, device="cuda").contiguous(memory_format=CL).requires_grad_(True)
    y1 = ew.Dropout2dFn.apply(d, 0.3)
    y2 = ew.Dropout2dFn.apply(d, 0.3)
    per_map = y1.detach().amax(dim=(2, 3))
    assert bool(((y1.detach() == per_map[..., None, None])).all())                       # constant per (sample, channel) map
    vals = torch.unique(per_map)
    assert len(vals) == 2 and float(vals[0]) == 0.0 and abs(float(vals[1]) - 1 / 0.7) < 1e-6
    rate = float((per_map == 0).float().mean())
    assert 0.27 < rate < 0.33, rate
    assert not torch.equal(y1, y2)
    y1.sum().backward()
    _close(d.grad, y1.detach(), 0.0)                                                    # d(sum y)/dx = mask


def test_fused_relu_and_skip_epilogues_match_the_unfused_graph():
    """conv2d(m, x, act=1, residual=skip) (one kernel forward, masked gradient backward) == relu(conv(x) + skip) through autograd."""
    from densefusion_b200.lib import conv_tc
    g = torch.Generator().manual_seed(9)
    m = torch.nn.Conv2d(64, 64, 3, padding=2, dilation=2, bias=False).cuda()
    x = _rand(2, 64, 20, 20, seed=11).requires_grad_(True)
    s = _rand(2, 64, 20, 20, seed=12).requires_grad_(True)
    y = conv_tc.conv2d(m, x, act=1, residual=s)
    gy = _rand(*y.shape, seed=13)
    y.backward(gy)
    got = (y.detach().clone(), x.grad.clone(), s.grad.clone(), m.weight.grad.clone())
    x.grad = s.grad = m.weight.grad = None
    yr = torch.relu(conv_tc.conv2d(m, x) + s)
    yr.backward(gy)
    _close(got[0], yr, 2e-6)
    _close(got[1], x.grad, 1e-5)
    _close(got[2], s.grad, 0.0)
    _close(got[3], m.weight.grad, 1e-5)


def test_estimator_training_step_launches_no_library_compute_kernels():
    """One estimator-phase forward + backward on the tensor-core training path: the profiler's kernel list holds no cuDNN / cuBLAS /
    CUTLASS kernel and, of ATen, only what autograd itself adds -- gradient accumulation (add), zero fills and layout copies."""
    from torch.profiler import ProfilerActivity, profile
    from densefusion_b200 import synth
    from densefusion_b200.lib.loss import Loss
    from util import build_nets
    est, _, _, _ = build_nets(500, 21, 3)
    est.train().requires_grad_(True)
    d = {k: v.cuda() for k, v in synth.synth_crop(3, 500, 500, 21, (80, 80), obj=4).items()}
    crit = Loss(500, synth.YCB_SYM)

    def step():
        r, t, c, _ = est(d["img"], d["points"], d["choose"], d["idx"])
        loss, *_ = crit(r, t, c, d["target"], d["model_points"], d["idx"], d["points"], 0.015, False)
        loss.backward()
    step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages() if e.device_time_total > 0]
    lib_kernels = [n for n in names if any(t in n.lower() for t in ("cudnn", "cublas", "cutlass", "sm90_", "sm100_", "gemv", "xmma"))]
    assert not lib_kernels, lib_kernels
    allowed = ("CUDAFunctor_add", "FillFunctor", "direct_copy", "Memcpy", "Memset", "copy_kernel", "CatArrayBatchedCopy")
    aten = [n for n in names if "at::native" in n and not any(a in n for a in allowed)]
    assert not aten, aten
