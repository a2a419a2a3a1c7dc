"""K1/K2 parity: GEMM building blocks, the fused dense-fusion head and the refiner against the oracle
(fp32 torch-CPU restatement of lib/network.py) and the reference's own outputs (tests/golden)."""
import numpy as np
import pytest
import torch

from conftest import golden
from densefusion_b200 import synth
from oracle import df_oracle as O
from util import build_nets, rel

pytestmark = pytest.mark.gpu

import os


def _variants(default):
    """DF_TEST_VARIANTS=5,6 restricts the tensor-core kernel variants under test (bring-up runs in separate processes)."""
    env = os.environ.get("DF_TEST_VARIANTS")
    return [int(v) for v in env.split(",")] if env else default


# stated bounds per arithmetic mode (max-abs error / max-abs value)
TOL = {"fp32": 1e-4, "3xtf32": 1e-4, "hybrid": 1e-4, "hybrid16": 1e-4, "hybrid16s": 1e-4, "tf32": 5e-3}


def _ref_linear(x, w, b, relu):
    y = x.double() @ w.double().t() + (0 if b is None else b.double())
    return torch.relu(y) if relu else y


@pytest.mark.parametrize("M,N,K", [(500, 64, 32), (1000, 128, 64), (4000, 512, 256), (777, 1920, 384),
                                   (8, 1920, 1024), (33, 1024, 1024), (129, 128, 16)])
def test_gemm_fp32_vs_float64(M, N, K):
    from densefusion_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    y = ops.linear(x.cuda(), w.cuda(), b.cuda(), relu=True)
    assert rel(y, _ref_linear(x, w, b, True)) < 2e-6


@pytest.mark.parametrize("variant", _variants([5, 6, 7]))
@pytest.mark.parametrize("precision", ["3xtf32", "tf32", "hybrid", "hybrid16", "hybrid16s"])
def test_gemm_tensor_core_variants(variant, precision):
    """variant 5: persistent kernel, one CTA per tile, A operand by TMA; 6 / 7: CTA pairs (cta_group::2, 256-row tiles) with
    2 / 4 TMEM A stages.  "hybrid16s" (32-column TMEM A stages) exists on variant 6, the production kernel."""
    from densefusion_b200 import ops
    g = torch.Generator().manual_seed(variant)
    M, N, K = 1000, 512, 384
    if precision == "hybrid16s" and variant != 6:
        pytest.skip("the two-plane fp16 arithmetic exists in the production kernel (variant 6) only")
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    ops.TC_VARIANT = variant
    try:
        y = ops.linear(x.cuda(), w.cuda(), b.cuda(), relu=False, precision=precision)
        torch.cuda.synchronize()
    finally:
        ops.TC_VARIANT = 0
    err = rel(y, _ref_linear(x, w, b, False))
    print(f"tc variant {variant} {precision}: rel err {err:.3e}")
    assert err < (3e-3 if precision == "tf32" else 1e-5)   # fp32 SIMT reaches ~1e-6; 3xTF32 measured 3e-6


def test_gemm_tensor_core_identity_layout():
    """A = structured, W = identity: any layout / swizzle / lane mix-up shows as a permutation (errors of order one); the values
    themselves only differ by the accumulator-truncation compensation (a factor 1 + O(1e-6))."""
    from densefusion_b200 import ops
    M, K = 256, 128
    x = (torch.arange(M * K, dtype=torch.float32).view(M, K) % 4093) / 64.0
    w = torch.eye(K)
    for variant in _variants([5, 6, 7]):
        ops.TC_VARIANT = variant
        try:
            y = ops.linear(x.cuda(), w.cuda(), None, precision="3xtf32")
            torch.cuda.synchronize()
        finally:
            ops.TC_VARIANT = 0
        bad = ((y.cpu() - x).abs() > 2e-6 * x.abs() + 1e-9).nonzero()
        assert bad.numel() == 0, f"variant {variant}: first bad index {bad[:4].tolist()}"


@pytest.mark.parametrize("M,N,K,groups", [(1000, 512, 384, 1), (64000 // 8, 1920, 384, 1), (3000, 256, 640, 3), (515, 128, 256, 3),
                                            (2500, 1024, 512, 1), (300, 64, 64, 1), (96, 1024, 512, 1), (700, 512, 4608, 1)])
def test_hybrid16s_all_tile_widths_and_groups(M, N, K, groups):
    """"hybrid16s" on every tile width the launcher picks (64 .. 256), grouped layers and a multi-run k loop, against float64 and
    within rounding of "hybrid16" (different planes, same 22-bit operands)."""
    from densefusion_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K * groups, generator=g).cuda()
    W = ops.SplitWeight((torch.randn(groups, N, K, generator=g) / K ** 0.5).cuda())
    b = torch.randn(N * groups, generator=g).cuda()

    def run(prec):
        C = torch.zeros(M, N * groups, device="cuda")
        ops.gemm(A, W, b, C, M=M, N=N, K=K, lda=K * groups, ldw=K, ldc=N * groups, relu=True, precision=prec, groups=groups,
                 a_gs=K, w_gs=N * K, bias_gs=N, c_gs=N)
        return C
    got, other = run("hybrid16s"), run("hybrid16")
    torch.cuda.synchronize()
    ref64 = torch.relu(torch.einsum("mgk,gnk->mgn", A.double().view(M, groups, K), W.w.double().view(groups, N, K)).reshape(M, -1) + b.double())
    assert rel(got, ref64) < 5e-6 and rel(got, other) < 1e-5


@pytest.mark.parametrize("precision", ["3xtf32", "hybrid", "hybrid16"])
@pytest.mark.parametrize("scale_a,scale_w", [(1e6, 1.0), (1.0, 1e6), (1e-7, 1.0), (1.0, 1e-7), (3e5, 1e-6), (1e3, 1.0), (1.0, 1e3),
                                             (1e-2, 1.0), (1.0, 1e-2), (1e3, 1e-2)])
def test_gemm_parity_modes_operand_range(precision, scale_a, scale_w):
    """Operand range of the fp32-parity modes, against float64, in the max-norm and element-wise (relative with an absolute
    floor of 1e-2 of the output scale: a dot product's absolute error scales with |a||w|, not with the individual result).
      * "3xtf32" / "hybrid": the main term keeps the fp32 exponent -> no range restriction (operands scaled by 1e6 / 1e-7 pass).
      * "hybrid16": the main term is fp16, so parity needs the operands that carry the dot product inside fp16's normal range,
        6.1e-5 <= |x| <= 65504 (smaller entries of an in-range tensor only lose bits that do not matter at the tensor's scale).
        Outside, the remainder x - fp16(x) is as large as x itself and only has bf16's 8 bits: the result degrades to a bf16-grade
        2^-8 (finite, saturating conversion) -- asserted here as the documented behaviour.  With BOTH operands out of range the
        dropped a_lo x w_lo term is no longer small and no accuracy is claimed (finite results only)."""
    from densefusion_b200 import ops
    from util import rel_elementwise
    g = torch.Generator().manual_seed(11)
    M, N, K = 777, 384, 512
    x = torch.randn(M, K, generator=g) * scale_a
    w = torch.randn(N, K, generator=g) / K ** 0.5 * scale_w
    b = torch.randn(N, generator=g) * (scale_a * scale_w)
    y = ops.linear(x.cuda(), w.cuda(), b.cuda(), relu=False, precision=precision)
    want = _ref_linear(x, w, b, False)
    err, err_el = rel(y, want), rel_elementwise(y, want, floor=1e-2)
    print(f"{precision} scale {scale_a:g} x {scale_w:g}: max-norm {err:.3e}, element-wise {err_el:.3e}")
    in_fp16_range = all(1e-3 <= s <= 1e4 for s in (scale_a, scale_w / K ** 0.5 * 20))
    both_out = not (1e-3 <= scale_a <= 1e4) and not (1e-3 <= scale_w <= 1e4)
    if precision.startswith("hybrid16") and not in_fp16_range:
        assert bool(torch.isfinite(y).all()) and (both_out or err < 2e-2)
    else:
        assert err < 1e-5 and err_el < 1e-3


def test_pack_f16s_planes_and_scale_bit_exact():
    """df_pack_f16s against a torch restatement: s = the power of two that brings the tensor's largest entry into [2^14, 2^15),
    planes = per row and k-block [fp16(w s) x32 | fp16(w s - fp16(w s)) x32], scale record (1/s, s)."""
    from densefusion_b200 import ops
    g = torch.Generator().manual_seed(4)
    for rows, K, mag in ((96, 64, 1.0), (257, 384, 3e-3), (64, 4608, 40.0), (8, 32, 1e-20), (8, 32, 0.0)):
        w = torch.randn(rows, K, generator=g) * mag
        w[0, 0] = 0.0
        planes, scale = ops.SplitWeight(w.cuda()).planes16s()
        torch.cuda.synchronize()
        m = float(w.abs().max())
        e = 14 - (int(torch.tensor(m).log2().floor()) if m > 0 else 14)
        if m > 0 and int(torch.tensor(m).log2().floor()) < -100:
            e = 0
        sc = 2.0 ** e
        assert float(scale[0]) == 1.0 / sc and float(scale[1]) == sc
        ws = w * sc
        hi = ws.half()
        lo = (ws - hi.float()).half()
        want = torch.cat([hi.view(rows, K // 32, 32), lo.view(rows, K // 32, 32)], 2).reshape(rows, 2 * K)
        got = planes.cpu().view(torch.float16).view(rows, 2 * K)
        assert torch.equal(got.view(torch.int16), want.view(torch.int16))
        if m > 1e-10:
            assert 2.0 ** 14 <= m * sc < 2.0 ** 15


@pytest.mark.parametrize("relu_input", [False, True])
@pytest.mark.parametrize("scale_a,scale_w", [(1.0, 1.0), (1e6, 1.0), (1.0, 1e6), (1e-7, 1.0), (1.0, 1e-7), (3e5, 1e-6), (1e-6, 3e5),
                                             (1e-12, 1e12), (1e-2, 1e-2)])
def test_hybrid16s_follows_the_operand_scales(relu_input, scale_a, scale_w):
    """"hybrid16s" carries both operands as two fp16 planes; the power-of-two scales (weights: from the tensor's maximum at pack
    time; activations: from the kernel's own 4096-element sample) make the arithmetic independent of the operands' magnitude --
    fp32 parity at every scale where "hybrid16" only reaches it inside fp16's normal range."""
    from densefusion_b200 import ops
    from util import rel_elementwise
    g = torch.Generator().manual_seed(17)
    M, N, K = 1500, 384, 1024
    x = torch.randn(M, K, generator=g) * scale_a
    if relu_input:
        x = torch.relu(x)
    w = torch.randn(N, K, generator=g) / K ** 0.5 * scale_w
    b = torch.randn(N, generator=g) * (scale_a * scale_w)
    y = ops.linear(x.cuda(), w.cuda(), b.cuda(), relu=False, precision="hybrid16s")
    want = _ref_linear(x, w, b, False)
    err, err_el = rel(y, want), rel_elementwise(y, want, floor=1e-2)
    print(f"hybrid16s scale {scale_a:g} x {scale_w:g} relu_input={relu_input}: max-norm {err:.3e}, element-wise {err_el:.3e}")
    assert err < 5e-6 and err_el < 1e-3


def test_hybrid16s_fixed_activation_scale_and_outliers():
    """What the sampled scale protects against, shown with FIXED scales (a_log2): an operand far below the scale's sweet spot
    keeps an absolute error of 2^-25 per entry (graceful: 1e-2-sized activations at scale 1 still reach 1e-5), one far above
    saturates.  And what the sample cannot see -- a single entry far larger than everything sampled -- stays exact up to
    ~2000x the sampled maximum (65504 / 2^5)."""
    from densefusion_b200 import ops
    g = torch.Generator().manual_seed(23)
    M, N, K = 1024, 256, 512
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()

    def err(x, a_log2=None):
        C = torch.empty(M, N, device="cuda")
        ops.gemm(x, ops.SplitWeight(w), None, C, M=M, N=N, K=K, lda=K, ldw=K, ldc=N, relu=False, precision="hybrid16s", a_log2=a_log2)
        ref = x.double() @ w.double().t()
        return float((C.double() - ref).abs().max() / ref.abs().max())

    x = torch.randn(M, K, generator=g).cuda()
    assert err(x) < 3e-6 and err(x, a_log2=0) < 3e-6 and err(x, a_log2=5) < 3e-6
    assert err(x * 1e-2, a_log2=0) < 1e-5                       # 2^-25 absolute on entries of 1e-2
    assert err(x * 1e-4, a_log2=0) > 1e-5                       # ... and that is where a fixed scale of 1 stops
    assert err(x * 1e-4) < 3e-6 and err(x * 1e4) < 3e-6         # the sampled scale follows the operand
    for big, ok in ((1e2, True), (1e3, True), (1.5e3, True)):
        x2 = x.clone()
        x2[777, 13] = big
        assert (err(x2) < 3e-6) == ok, big
    # an "outlier channel": one column 300x larger than all the others (the diagonal sample visits every 4-column group)
    x3 = x.clone()
    x3[:, 301] *= 300.0
    assert err(x3) < 3e-6


@pytest.mark.parametrize("variant", _variants([5, 6, 7]))
def test_gemm_tensor_core_epilogues_match_fp32_kernel(variant):
    """per-crop bias, grouped (block-diagonal) and column-pool epilogues: 3xtf32 kernel vs the exact-fp32 kernel."""
    from densefusion_b200 import ops
    g = torch.Generator().manual_seed(77)
    crops, n = 5, 500
    rows = crops * n

    def run(precision, fn):
        ops.TC_VARIANT = variant
        try:
            return fn(precision)
        finally:
            ops.TC_VARIANT = 0

    # per-crop bias
    A = torch.randn(rows, 384, generator=g).cuda()
    W = ops.SplitWeight((torch.randn(1920, 384, generator=g) / 20).cuda())
    bias = torch.randn(crops, 1920, generator=g).cuda()

    def percrop(prec):
        C = torch.empty(rows, 1920, device="cuda")
        ops.gemm(A, W, bias, C, M=rows, N=1920, K=384, lda=384, ldw=384, ldc=1920, relu=True, precision=prec,
                 bias_crop_stride=1920, rows_per_crop=n)
        return C
    assert rel(run("3xtf32", percrop), percrop("fp32")) < 1e-5
    # grouped towers
    A3 = torch.randn(rows, 1920, generator=g).cuda()
    W3 = ops.SplitWeight((torch.randn(3, 256, 640, generator=g) / 25).cuda())
    b3 = torch.randn(768, generator=g).cuda()

    def grouped(prec):
        C = torch.zeros(rows, 768, device="cuda")
        ops.gemm(A3, W3, b3, C, M=rows, N=256, K=640, lda=1920, ldw=640, ldc=768, relu=True, precision=prec, groups=3,
                 a_gs=640, w_gs=256 * 640, bias_gs=256, c_gs=256)
        return C
    assert rel(run("3xtf32", grouped), grouped("fp32")) < 1e-5
    # pooled (crop-aligned tiles, 500 = 3*128 + 116 rows)
    A6 = torch.randn(rows, 512, generator=g).cuda()
    W6 = ops.SplitWeight((torch.randn(1024, 512, generator=g) / 22).cuda())
    b6 = torch.randn(1024, generator=g).cuda()

    def pooled(prec):
        part = torch.zeros(crops, 4, 1024, device="cuda")
        ops.gemm(A6, W6, b6, None, M=rows, N=1024, K=512, lda=512, ldw=512, ldc=0, relu=True, precision=prec,
                 rows_per_crop=n, pool_partial=part)
        return part.sum(1)
    got, want = run("3xtf32", pooled), pooled("fp32")
    ref64 = torch.relu(A6.double() @ W6.w.double().t() + b6.double()).view(crops, n, 1024).sum(1)
    assert rel(got, ref64) < 1e-5 and rel(want, ref64) < 1e-5
    # third tower layer: three groups of N = 128 (64 weight rows per CTA in the paired kernel), A is a column slice
    A4 = torch.randn(rows, 768, generator=g).cuda()
    W4 = ops.SplitWeight((torch.randn(3, 128, 256, generator=g) / 16).cuda())
    b4 = torch.randn(384, generator=g).cuda()

    def tower3(prec):
        C = torch.zeros(rows, 384, device="cuda")
        ops.gemm(A4, W4, b4, C, M=rows, N=128, K=256, lda=768, ldw=256, ldc=384, relu=True, precision=prec, groups=3,
                 a_gs=256, w_gs=128 * 256, bias_gs=128, c_gs=128)
        return C
    assert rel(run("3xtf32", tower3), tower3("fp32")) < 1e-5
    # conv5 of PoseNetFeat: A = columns 128:384 of the 384-wide point-feature buffer, M not a multiple of 256
    pf = torch.randn(rows - 37, 384, generator=g).cuda()
    W5 = ops.SplitWeight((torch.randn(512, 256, generator=g) / 16).cuda())
    b5 = torch.randn(512, generator=g).cuda()

    def conv5(prec):
        C = torch.zeros(rows - 37, 512, device="cuda")
        ops.gemm(pf[:, 128:], W5, b5, C, M=rows - 37, N=512, K=256, lda=384, ldw=256, ldc=512, relu=True, precision=prec)
        return C
    assert rel(run("3xtf32", conv5), conv5("fp32")) < 1e-5


def test_per_crop_bias_is_not_read_past_its_last_row():
    """Regression for the round-1 'ConvS2Fn' illegal address: with M = crops * 500 not a multiple of the 256-row tile, the last
    warp of the last tile lies entirely in the masked tail and used to load the bias vector of crop index `crops` -- one row past
    the (crops, N) buffer.  The buffer here ends exactly at the end of its own device allocation, so the stray read leaves the
    mapping (a fault) instead of landing in a neighbouring tensor; the result must also equal the exact-fp32 kernel."""
    from densefusion_b200 import ops
    g = torch.Generator().manual_seed(5)
    for crops in (6, 4, 1):
        n = 500
        rows = crops * n
        big = torch.empty(12 * 1024 * 1024 // 4, device="cuda")                # > 10 MB: a device allocation of its own
        bias = big[-crops * 1920:].view(crops, 1920)
        bias.copy_(torch.randn(crops, 1920, generator=g))
        A = torch.randn(rows, 384, generator=g).cuda()
        W = ops.SplitWeight((torch.randn(1920, 384, generator=g) / 20).cuda())

        def run(prec):
            C = torch.empty(rows, 1920, device="cuda")
            ops.gemm(A, W, bias, C, M=rows, N=1920, K=384, lda=384, ldw=384, ldc=1920, relu=True, precision=prec,
                     bias_crop_stride=1920, rows_per_crop=n)
            torch.cuda.synchronize()
            return C
        assert rel(run("hybrid16"), run("fp32")) < 1e-5
        del big, bias


@pytest.mark.parametrize("precision", ["fp32", "3xtf32", "tf32", "hybrid", "hybrid16", "hybrid16s"])
@pytest.mark.parametrize("n,o,B", [(500, 21, 3), (1000, 13, 2)])
def test_head_vs_oracle(precision, n, o, B):
    est, _, est_sd, _ = build_nets(n, o, seed=5)
    est.precision = precision
    g = torch.Generator().manual_seed(n + B)
    x = torch.randn(B, n, 3, generator=g) * 0.05 + torch.tensor([0.0, 0.0, 0.8])
    emb = torch.cat([synth.synth_embedding(100 + i, n) for i in range(B)], 0)            # (B,32,N)
    obj = torch.randint(0, o, (B, 1), generator=g)
    emb_pm = emb.permute(0, 2, 1).reshape(B * n, 32).contiguous()
    r, t, c = est.head(x.cuda(), emb_pm.cuda(), obj.cuda())
    worst = 0.0
    for b in range(B):
        with torch.no_grad():
            rr, tt, cc = O.posenet_head(est_sd, x[b:b + 1], emb[b:b + 1], obj[b:b + 1], o)
        errs = (rel(r[b], rr[0]), rel(t[b], tt[0]), rel(c[b], cc[0]))
        worst = max(worst, *errs)
        assert max(errs) < TOL[precision], f"{precision} crop {b}: {errs}"
        if precision != "tf32":
            margin = torch.sort(cc.view(-1), descending=True)[0]
            if float(margin[0] - margin[1]) > 1e-5:
                assert int(torch.argmax(c[b].view(-1))) == int(torch.argmax(cc.view(-1)))
    print(f"head {precision} n={n}: worst rel err {worst:.3e}")


@pytest.mark.parametrize("precision", ["fp32", "3xtf32", "hybrid", "hybrid16", "hybrid16s"])
def test_refiner_vs_oracle(precision):
    n, o, B = 500, 21, 4
    _, ref, _, ref_sd = build_nets(n, o, seed=6)
    ref.precision = precision
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, n, 3, generator=g) * 0.05
    emb = torch.cat([synth.synth_embedding(200 + i, n) for i in range(B)], 0)
    obj = torch.randint(0, o, (B, 1), generator=g)
    emb_pm = emb.permute(0, 2, 1).reshape(B * n, 32).contiguous()
    r, t = ref.refine(x.cuda(), emb_pm.cuda(), obj.cuda())
    for b in range(B):
        with torch.no_grad():
            rr, tt = O.refiner_forward(ref_sd, x[b:b + 1], emb[b:b + 1], obj[b:b + 1], o)
        assert rel(r[b], rr[0]) < TOL[precision] and rel(t[b], tt[0]) < TOL[precision]
    # reference-contract forward (bs=1, channel-major emb)
    r1, t1 = ref(x[0:1].cuda(), emb[0:1].cuda(), obj[0:1].cuda())
    assert tuple(r1.shape) == (1, 4) and tuple(t1.shape) == (1, 3)
    assert torch.equal(r1[0], r[0]) or rel(r1[0], r[0]) < 1e-6


@pytest.mark.parametrize("name", ["c0_linemod_add", "c1_ycb_adds"])
def test_posenet_dropin_vs_reference_golden(name):
    """Full drop-in PoseNet.forward (torch/cuDNN encoder in fp32 + fused head) on the golden inputs."""
    g = golden(name)
    case, n, o, m, h, w, obj, seed, iters = [int(v) for v in g["meta"]]
    est, ref, _, _ = build_nets(n, o, seed)
    d = {k: v.cuda() for k, v in synth.synth_crop(case, n, m, o, (h, w), obj).items()}
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        r, t, c, emb = est(d["img"], d["points"], d["choose"], d["idx"])
    assert tuple(r.shape) == (1, n, 4) and tuple(t.shape) == (1, n, 3) and tuple(c.shape) == (1, n, 1)
    assert tuple(emb.shape) == (1, 32, n) and not emb.requires_grad
    assert rel(emb, g["emb"]) < 1e-4
    assert rel(r, g["pred_r"]) < 1e-4 and rel(t, g["pred_t"]) < 1e-4 and rel(c, g["pred_c"]) < 1e-4
    assert int(torch.argmax(c.view(-1))) == int(g["which_max"])
    # refiner on the reference's own training-chain inputs
    emb_ref = torch.from_numpy(g["emb"]).cuda()
    pts = torch.from_numpy(g["new_points"]).cuda()
    with torch.no_grad():
        rr, tt = ref(pts, emb_ref, d["idx"])
    assert rel(rr, g["train_r0"]) < 1e-4 and rel(tt, g["train_t0"]) < 1e-4


def test_point_major_entry_points_are_inference_only():
    from densefusion_b200.lib.network import PoseRefineNet
    net = PoseRefineNet(500, 13).cuda()
    with pytest.raises(RuntimeError):
        net.refine(torch.zeros(1, 500, 3, device="cuda"), torch.zeros(500, 32, device="cuda"),
                   torch.zeros(1, dtype=torch.long, device="cuda"))


@pytest.mark.parametrize("shape,size,align", [((3, 512, 10, 10), (10, 10), False), ((2, 512, 1, 1), (15, 15), False),
                                               ((2, 64, 6, 6), (20, 20), False), ((3, 1024, 10, 10), (20, 20), True),
                                               ((2, 64, 40, 40), (80, 80), True), ((1, 7, 5, 9), (10, 18), True)])
def test_upsample_bilinear_matches_torch(shape, size, align):
    from densefusion_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g)
    want = torch.nn.functional.interpolate(x, size=size, mode="bilinear", align_corners=align)
    got = ops.upsample_bilinear(x.cuda(), size, align)
    assert tuple(got.shape) == tuple(want.shape)
    assert rel(got, want) < 2e-6
