"""Training-mode forward / backward of the dense-fusion head (K1) and the refiner (K2) as autograd Functions over
the C-ABI kernels -- what `loss.backward()` / `dis.backward()` do in tools/train.py:152-161, without any torch op on
the activation path.

Arithmetic: forward GEMMs in `PRECISION` (default "hybrid16": tcgen05, fp32 parity for activation-scale operands), data
gradients in `GRAD_PRECISION` (default "hybrid": the A operand there is a gradient, routinely 1e-6 .. 1e-9 in magnitude, which
fp16's exponent range cannot carry -- the TF32 main term of "hybrid" keeps the fp32 exponent), weight gradients as 3xTF32 GEMMs
over the rows (df_conv_wgrad_tc) or exact fp32; PRECISION = "fp32" selects the exact FFMA kernels everywhere.

Forward keeps every ReLU output (pf, h5, h6, h1, h2, h3): the backward of ReLU only needs `output > 0`, which is
fused into the data-gradient GEMM epilogue (`relu_mask`).  Weight gradients are split over the rows and reduced in a
fixed order, so a training step is run-to-run deterministic except for the embedding scatter-add when `choose` holds
duplicate pixels (atomicAdd)."""
from __future__ import annotations

from typing import List

import torch

from . import ops
from ._C import check, lib, ptr, stream

HEAD_PARAMS = (["feat.conv1", "feat.e_conv1", "feat.conv2", "feat.e_conv2", "feat.conv5", "feat.conv6"]
               + [f"conv{l}_{b}" for l in (1, 2, 3, 4) for b in "rtc"])
REFINER_PARAMS = (["feat.conv1", "feat.e_conv1", "feat.conv2", "feat.e_conv2", "feat.conv5", "feat.conv6"]
                  + [f"conv{l}_{b}" for l in (1, 2, 3) for b in "rt"])


def _get(net, dotted):
    m = net
    for part in dotted.split("."):
        m = getattr(m, part)
    return m


def param_list(net, names) -> List[torch.nn.Parameter]:
    out = []
    for n in names:
        m = _get(net, n)
        out += [m.weight, m.bias]
    return out


def _f(*shape, dev):
    return torch.empty(*shape, device=dev, dtype=torch.float32)


def _w2(p):
    return p.detach().reshape(p.shape[0], -1).float().contiguous()


# Arithmetic of the forward and data-gradient GEMMs: an fp32-parity tensor-core mode ("hybrid16" / "hybrid" / "3xtf32") or "fp32"
# (exact FFMA).  Weight gradients (reductions over the rows) and everything too small for a 128-row MMA tile are always
# exact fp32.
PRECISION = "hybrid16"
GRAD_PRECISION = "hybrid"       # data-gradient GEMMs (A operand = a gradient): an exponent-safe parity mode, see the module docstring


def grad_precision() -> str:
    return "fp32" if PRECISION == "fp32" else (GRAD_PRECISION if PRECISION.startswith("hybrid16") else PRECISION)


# ---- thin kernel wrappers ----------------------------------------------------------------------------------
def _gemm(A, W, bias, C, **kw):
    if PRECISION != "fp32" and ops.tc_eligible(kw["M"], kw["N"], kw["K"]):
        ops.gemm(A, ops.SplitWeight(W), bias, C, precision=PRECISION, short_runs=True, **kw)
    else:
        ops.gemm(A, W, bias, C, precision="fp32", **kw)


WGRAD_TC = True         # weight gradients of the wide layers on df_conv_wgrad_tc (3xTF32 GEMM over the rows)


def _wgrad(dY, ldy, X, ldx, M, N, K, groups=1, dy_gs=0, x_gs=0):
    if WGRAD_TC and PRECISION != "fp32" and K % 64 == 0 and N % 4 == 0 and M >= 256:
        # dW[g] (N,K) = dY[:, g]^T X[:, g]: the rows are the "pixels" of a 1x1 convolution's weight gradient
        out = _f(groups, N, K, dev=dY.device)
        n = int(lib.df_conv_wgrad_scratch_floats(1, 1, M, K, N, 1, 1))
        scratch = _f(n, dev=dY.device)
        for g in range(groups):
            check(lib.df_conv_wgrad_tc(X.data_ptr() + 4 * g * x_gs, ldx, dY.data_ptr() + 4 * g * dy_gs, ldy, 1, 1, M, K, N, 1, 1,
                                       ptr(scratch), ptr(out[g]), stream()), "df_conv_wgrad_tc")
        return out
    splits = max(1, min(32, M // 256))
    part = _f(splits, groups, N, K, dev=dY.device)
    check(lib.df_gemm_wgrad_fp32(ptr(dY), ldy, ptr(X), ldx, ptr(part), M, N, K, groups, splits, dy_gs, x_gs, stream()),
          "df_gemm_wgrad_fp32")
    out = _f(groups, N, K, dev=dY.device)
    check(lib.df_reduce_partials(ptr(part), splits, groups * N * K, ptr(out), 0, stream()), "df_reduce_partials")
    return out


def _dgrad(dY, ldy, Wt, ldw, dX, ldx, M, N, K, groups=1, dy_gs=0, w_gs=0, dx_gs=0, mask=None, accumulate=False):
    if (PRECISION != "fp32" and not accumulate and ops.tc_eligible(M, N, K) and (groups == 1 or (w_gs == N * ldw and dx_gs == N))
            and Wt.is_contiguous()):
        # tensor-core data gradient; the ReLU mask of the layer below is applied by a separate element-wise pass
        ops.gemm(dY, ops.SplitWeight(Wt), None, dX, M=M, N=N, K=K, lda=ldy, ldw=ldw, ldc=ldx, relu=False, precision=grad_precision(),
                 groups=groups, a_gs=dy_gs, w_gs=w_gs, c_gs=dx_gs, short_runs=True)
        if mask is not None:
            _mask_inplace(dX, mask, ldx, N * groups, M)
        return
    check(lib.df_gemm_dgrad_fp32(ptr(dY), ldy, ptr(Wt), ldw, ptr(dX), ldx, M, N, K, groups, dy_gs, w_gs, dx_gs, ptr(mask),
                                 1 if accumulate else 0, stream()), "df_gemm_dgrad_fp32")


def _colsum(X, ldx, rows_per_group, groups, C):
    out = _f(groups, C, dev=X.device)
    check(lib.df_colsum_rows(ptr(X), ldx, rows_per_group, groups, C, ptr(out), 0, stream()), "df_colsum_rows")
    return out


def _mask_inplace(d, act, ld, cols, rows):
    check(lib.df_relu_mask_inplace(ptr(d), ptr(act), ld, cols, rows, stream()), "df_relu_mask_inplace")


# ---- shared point-feature trunk (PoseNetFeat / PoseRefineNetFeat) -----------------------------------------------
class _FeatState:
    pass


def _feat_forward(net_feat, x, emb_pm, B, n, full_conv5: bool) -> _FeatState:
    """x (rows,3), emb_pm (rows,32).  Keeps pf / h5 / h6 for the backward; g = per-crop mean of h6."""
    dev, rows = x.device, B * n
    st = _FeatState()
    st.x, st.emb, st.B, st.n, st.full5 = x, emb_pm, B, n, full_conv5
    w = {k: _w2(getattr(net_feat, k).weight) for k in ("conv1", "e_conv1", "conv2", "e_conv2", "conv5", "conv6")}
    b = {k: getattr(net_feat, k).bias.detach().float().contiguous() for k in w}
    st.w = w
    pf = _f(rows, 384, dev=dev)
    check(lib.df_xyz_conv(ptr(x), ptr(w["conv1"]), ptr(b["conv1"]), ptr(pf), 384, rows, stream()), "df_xyz_conv")
    _gemm(emb_pm, w["e_conv1"], b["e_conv1"], pf[:, 64:], M=rows, N=64, K=32, lda=32, ldw=32, ldc=384, relu=True)
    _gemm(pf, w["conv2"], b["conv2"], pf[:, 128:], M=rows, N=128, K=64, lda=384, ldw=64, ldc=384, relu=True)
    _gemm(pf[:, 64:], w["e_conv2"], b["e_conv2"], pf[:, 256:], M=rows, N=128, K=64, lda=384, ldw=64, ldc=384, relu=True)
    k5 = 384 if full_conv5 else 256
    a5 = pf if full_conv5 else pf[:, 128:]
    h5 = _f(rows, 512, dev=dev)
    _gemm(a5, w["conv5"], b["conv5"], h5, M=rows, N=512, K=k5, lda=384, ldw=k5, ldc=512, relu=True)
    h6 = _f(rows, 1024, dev=dev)
    _gemm(h5, w["conv6"], b["conv6"], h6, M=rows, N=1024, K=512, lda=512, ldw=512, ldc=1024, relu=True)
    gsum = _colsum(h6, 1024, n, B, 1024)
    g = _f(B, 1024, dev=dev)
    check(lib.df_pool_finish(ptr(gsum), ptr(g), B, 1, 1024, n, stream()), "df_pool_finish")
    st.pf, st.h5, st.h6, st.g = pf, h5, h6, g
    return st


def _feat_backward(st: _FeatState, dg, dpf, need_demb: bool):
    """dg (B,1024) gradient of the pooled feature; dpf (rows,384) gradient already flowing into pf (or None).
    Returns ({name: (dW, db)}, demb or None)."""
    B, n, rows, dev = st.B, st.n, st.B * st.n, dg.device
    w, pf, h5, h6 = st.w, st.pf, st.h5, st.h6
    grads = {}
    dh6 = _f(rows, 1024, dev=dev)
    check(lib.df_pool_backward(ptr(dg), ptr(h6), ptr(dh6), n, 1024, rows, stream()), "df_pool_backward")
    grads["conv6"] = (_wgrad(dh6, 1024, h5, 512, rows, 1024, 512)[0], _colsum(dh6, 1024, rows, 1, 1024)[0])
    dh5 = _f(rows, 512, dev=dev)
    _dgrad(dh6, 1024, w["conv6"].t().contiguous(), 1024, dh5, 512, rows, 512, 1024, mask=h5)
    k5 = 384 if st.full5 else 256
    a5 = pf if st.full5 else pf[:, 128:]
    grads["conv5"] = (_wgrad(dh5, 512, a5, 384, rows, 512, k5)[0], _colsum(dh5, 512, rows, 1, 512)[0])
    w5t = w["conv5"].t().contiguous()                                   # (k5, 512)
    if dpf is None:
        dpf = torch.zeros(rows, 384, device=dev, dtype=torch.float32)
    _dgrad(dh5, 512, w5t, 512, dpf if st.full5 else dpf[:, 128:], 384, rows, k5, 512, accumulate=True)
    # second-level features x2 | e2 (columns 128:384): through their ReLU, then into x1 | e1
    _mask_inplace(dpf[:, 128:], pf[:, 128:], 384, 256, rows)
    grads["conv2"] = (_wgrad(dpf[:, 128:], 384, pf, 384, rows, 128, 64)[0], _colsum(dpf[:, 128:], 384, rows, 1, 128)[0])
    grads["e_conv2"] = (_wgrad(dpf[:, 256:], 384, pf[:, 64:], 384, rows, 128, 64)[0], _colsum(dpf[:, 256:], 384, rows, 1, 128)[0])
    _dgrad(dpf[:, 128:], 384, w["conv2"].t().contiguous(), 128, dpf, 384, rows, 64, 128, accumulate=True)
    _dgrad(dpf[:, 256:], 384, w["e_conv2"].t().contiguous(), 128, dpf[:, 64:], 384, rows, 64, 128, accumulate=True)
    # first-level features x1 | e1 (columns 0:128)
    _mask_inplace(dpf, pf, 384, 128, rows)
    grads["conv1"] = (_wgrad(dpf, 384, st.x, 3, rows, 64, 3)[0], _colsum(dpf, 384, rows, 1, 64)[0])
    grads["e_conv1"] = (_wgrad(dpf[:, 64:], 384, st.emb, 32, rows, 64, 32)[0], _colsum(dpf[:, 64:], 384, rows, 1, 64)[0])
    demb = None
    if need_demb:
        demb = _f(rows, 32, dev=dev)
        _dgrad(dpf[:, 64:], 384, w["e_conv1"].t().contiguous(), 64, demb, 32, rows, 32, 64)
    return grads, demb


def _select_backward(g_r, g_t, g_c, out_c, h, ldh, Wr, Wt, Wc, obj, rows_per_crop, num_obj, rows):
    dev = h.device
    crops = rows // rows_per_crop
    dh = _f(rows, ldh, dev=dev)
    gz, blk, bsum = _f(rows, 8, dev=dev), _f(crops, 8, 8, 128, dev=dev), _f(crops, 8, 8, dev=dev)     # DF_SELECT_SPLITS = 8
    z = lambda t: torch.zeros_like(t)
    dWr, dWt = z(Wr), z(Wt)
    dbr = torch.zeros(Wr.shape[0], device=dev)
    dbt = torch.zeros(Wt.shape[0], device=dev)
    dWc = z(Wc) if Wc is not None else None
    dbc = torch.zeros(Wc.shape[0], device=dev) if Wc is not None else None
    check(lib.df_select_out_backward(ptr(g_r), ptr(g_t), ptr(g_c), ptr(out_c), ptr(h), ldh, ptr(Wr), ptr(Wt), ptr(Wc),
                                     ptr(obj), rows_per_crop, num_obj, rows, ptr(dh), ptr(gz), ptr(blk), ptr(bsum),
                                     ptr(dWr), ptr(dbr), ptr(dWt), ptr(dbt), ptr(dWc), ptr(dbc), stream()),
          "df_select_out_backward")
    return dh, (dWr, dbr), (dWt, dbt), (dWc, dbc)


def _gc(t):
    return None if t is None else ops.f32c(t)


# ---- PoseNet head -----------------------------------------------------------------------------------------------
class PoseNetHeadFn(torch.autograd.Function):
    """(feature map (B,32,H,W), x (B,N,3), choose (B,1,N), obj (B,), net, *HEAD params) -> out_r, out_t, out_c, emb_cm."""

    @staticmethod
    def forward(ctx, feat, x, choose, obj, net, *params):
        B, n = x.shape[0], x.shape[1]
        rows, dev = B * n, x.device
        emb_pm, emb_cm = ops.gather_embedding(feat.detach(), choose)
        x2 = ops.f32c(x.detach()).view(rows, 3)
        obj = ops.i64c(obj).view(-1)
        st = _feat_forward(net.feat, x2, emb_pm, B, n, full_conv5=False)
        t1 = [_w2(getattr(net, f"conv1_{b}").weight) for b in "rtc"]
        w1l = torch.cat([w[:, :384] for w in t1], 0).contiguous()
        w1g = torch.cat([w[:, 384:] for w in t1], 0).contiguous()
        cat_b = lambda l: torch.cat([getattr(net, f"conv{l}_{b}").bias.detach().float() for b in "rtc"]).contiguous()
        w2 = torch.stack([_w2(getattr(net, f"conv2_{b}").weight) for b in "rtc"]).contiguous()
        w3 = torch.stack([_w2(getattr(net, f"conv3_{b}").weight) for b in "rtc"]).contiguous()
        w4 = [_w2(getattr(net, f"conv4_{b}").weight) for b in "rtc"]
        b4 = [getattr(net, f"conv4_{b}").bias.detach().float().contiguous() for b in "rtc"]
        gbias = _f(B, 1920, dev=dev)
        _gemm(st.g, w1g, cat_b(1), gbias, M=B, N=1920, K=1024, lda=1024, ldw=1024, ldc=1920, relu=False)
        h1 = _f(rows, 1920, dev=dev)
        _gemm(st.pf, w1l, gbias, h1, M=rows, N=1920, K=384, lda=384, ldw=384, ldc=1920, relu=True,
              bias_crop_stride=1920, rows_per_crop=n)
        h2 = _f(rows, 768, dev=dev)
        _gemm(h1, w2, cat_b(2), h2, M=rows, N=256, K=640, lda=1920, ldw=640, ldc=768, relu=True, groups=3, a_gs=640,
              w_gs=256 * 640, bias_gs=256, c_gs=256)
        h3 = _f(rows, 384, dev=dev)
        _gemm(h2, w3, cat_b(3), h3, M=rows, N=128, K=256, lda=768, ldw=256, ldc=384, relu=True, groups=3, a_gs=256,
              w_gs=128 * 256, bias_gs=128, c_gs=128)
        out_r, out_t, out_c = _f(B, n, 4, dev=dev), _f(B, n, 3, dev=dev), _f(B, n, 1, dev=dev)
        check(lib.df_select_out(ptr(h3), 384, ptr(w4[0]), ptr(b4[0]), ptr(w4[1]), ptr(b4[1]), ptr(w4[2]), ptr(b4[2]),
                                ptr(obj), n, net.num_obj, rows, ptr(out_r), ptr(out_t), ptr(out_c), stream()), "df_select_out")
        ctx.st, ctx.net = st, net
        ctx.saved = (w1l, w1g, w2, w3, w4, h1, h2, h3, out_c, obj, choose, feat.shape, feat.stride() if feat.is_contiguous() else None)
        ctx.mark_non_differentiable(emb_cm)
        return out_r, out_t, out_c, emb_cm

    @staticmethod
    def backward(ctx, g_r, g_t, g_c, _g_emb):
        st, net = ctx.st, ctx.net
        w1l, w1g, w2, w3, w4, h1, h2, h3, out_c, obj, choose, fshape, _ = ctx.saved
        B, n = st.B, st.n
        rows, dev = B * n, h1.device
        dh3, g4r, g4t, g4c = _select_backward(_gc(g_r), _gc(g_t), _gc(g_c), out_c, h3, 384, w4[0], w4[1], w4[2], obj, n,
                                              net.num_obj, rows)
        # tower layer 3
        db3 = _colsum(dh3, 384, rows, 1, 384)[0]
        dW3 = _wgrad(dh3, 384, h2, 768, rows, 128, 256, groups=3, dy_gs=128, x_gs=256)
        dh2 = _f(rows, 768, dev=dev)
        _dgrad(dh3, 384, w3.transpose(1, 2).contiguous(), 128, dh2, 768, rows, 256, 128, groups=3, dy_gs=128,
               w_gs=256 * 128, dx_gs=256, mask=h2)
        # tower layer 2
        db2 = _colsum(dh2, 768, rows, 1, 768)[0]
        dW2 = _wgrad(dh2, 768, h1, 1920, rows, 256, 640, groups=3, dy_gs=256, x_gs=640)
        dh1 = _f(rows, 1920, dev=dev)
        _dgrad(dh2, 768, w2.transpose(1, 2).contiguous(), 256, dh1, 1920, rows, 640, 256, groups=3, dy_gs=256,
               w_gs=640 * 256, dx_gs=640, mask=h1)
        # tower layer 1 = local 384 channels + per-crop bias carrying the global feature
        dgbias = _colsum(dh1, 1920, n, B, 1920)                                   # (B,1920)
        db1 = _colsum(dgbias, 1920, B, 1, 1920)[0]
        dW1l = _wgrad(dh1, 1920, st.pf, 384, rows, 1920, 384)[0]
        dW1g = _wgrad(dgbias, 1920, st.g, 1024, B, 1920, 1024)[0]
        dg = _f(B, 1024, dev=dev)
        _dgrad(dgbias, 1920, w1g.t().contiguous(), 1920, dg, 1024, B, 1024, 1920)
        dpf = _f(rows, 384, dev=dev)
        _dgrad(dh1, 1920, w1l.t().contiguous(), 1920, dpf, 384, rows, 384, 1920)
        need_feat = ctx.needs_input_grad[0]
        fg, demb = _feat_backward(st, dg, dpf, need_feat)
        dfeat = None
        if need_feat:
            dfeat = torch.zeros(fshape, device=dev, dtype=torch.float32)
            sb, sc, sh, sw = dfeat.stride()
            check(lib.df_gather_embedding_backward(ptr(demb), ptr(ops.i64c(choose).view(B, -1)), ptr(dfeat), sb, sc, sw, B, n,
                                                   fshape[2] * fshape[3], stream()), "df_gather_embedding_backward")
        out = []
        for name in HEAD_PARAMS:
            if name.startswith("feat."):
                dW, db = fg[name[5:]]
                out += [dW.view_as(_get(net, name).weight), db]
            else:
                layer, br = int(name[4]), "rtc".index(name[6])
                if layer == 1:
                    dW = torch.cat([dW1l[640 * br:640 * (br + 1)], dW1g[640 * br:640 * (br + 1)]], 1)
                    out += [dW.unsqueeze(-1).contiguous(), db1[640 * br:640 * (br + 1)].contiguous()]
                elif layer == 2:
                    out += [dW2[br].unsqueeze(-1).contiguous(), db2[256 * br:256 * (br + 1)].contiguous()]
                elif layer == 3:
                    out += [dW3[br].unsqueeze(-1).contiguous(), db3[128 * br:128 * (br + 1)].contiguous()]
                else:
                    dW, db = (g4r, g4t, g4c)[br]
                    out += [dW.unsqueeze(-1).contiguous(), db]
        return (dfeat, None, None, None, None, *out)


# ---- refiner --------------------------------------------------------------------------------------------------
class RefinerFn(torch.autograd.Function):
    """(x (B,N,3), emb_pm (B*N,32), obj (B,), net, *REFINER params) -> out_r (B,4), out_t (B,3)."""

    @staticmethod
    def forward(ctx, x, emb_pm, obj, net, *params):
        B, n = x.shape[0], x.shape[1]
        rows, dev = B * n, x.device
        obj = ops.i64c(obj).view(-1)
        st = _feat_forward(net.feat, ops.f32c(x.detach()).view(rows, 3), ops.f32c(emb_pm.detach()), B, n, full_conv5=True)
        w1 = torch.cat([_w2(net.conv1_r.weight), _w2(net.conv1_t.weight)], 0).contiguous()
        b1 = torch.cat([net.conv1_r.bias.detach().float(), net.conv1_t.bias.detach().float()]).contiguous()
        w2 = torch.stack([_w2(net.conv2_r.weight), _w2(net.conv2_t.weight)]).contiguous()
        b2 = torch.cat([net.conv2_r.bias.detach().float(), net.conv2_t.bias.detach().float()]).contiguous()
        w3 = [_w2(net.conv3_r.weight), _w2(net.conv3_t.weight)]
        b3 = [net.conv3_r.bias.detach().float().contiguous(), net.conv3_t.bias.detach().float().contiguous()]
        h1 = _f(B, 1024, dev=dev)
        _gemm(st.g, w1, b1, h1, M=B, N=1024, K=1024, lda=1024, ldw=1024, ldc=1024, relu=True)
        h2 = _f(B, 256, dev=dev)
        _gemm(h1, w2, b2, h2, M=B, N=128, K=512, lda=1024, ldw=512, ldc=256, relu=True, groups=2, a_gs=512,
              w_gs=128 * 512, bias_gs=128, c_gs=128)
        out_r, out_t = _f(B, 4, dev=dev), _f(B, 3, dev=dev)
        check(lib.df_select_out(ptr(h2), 256, ptr(w3[0]), ptr(b3[0]), ptr(w3[1]), ptr(b3[1]), None, None, ptr(obj), 1,
                                net.num_obj, B, ptr(out_r), ptr(out_t), None, stream()), "df_select_out")
        ctx.st, ctx.net = st, net
        ctx.saved = (w1, w2, w3, h1, h2, obj)
        return out_r, out_t

    @staticmethod
    def backward(ctx, g_r, g_t):
        st, net = ctx.st, ctx.net
        w1, w2, w3, h1, h2, obj = ctx.saved
        B, dev = st.B, h1.device
        dh2, g3r, g3t, _ = _select_backward(_gc(g_r), _gc(g_t), None, None, h2, 256, w3[0], w3[1], None, obj, 1,
                                            net.num_obj, B)
        db2 = _colsum(dh2, 256, B, 1, 256)[0]
        dW2 = _wgrad(dh2, 256, h1, 1024, B, 128, 512, groups=2, dy_gs=128, x_gs=512)
        dh1 = _f(B, 1024, dev=dev)
        _dgrad(dh2, 256, w2.transpose(1, 2).contiguous(), 128, dh1, 1024, B, 512, 128, groups=2, dy_gs=128,
               w_gs=512 * 128, dx_gs=512, mask=h1)
        db1 = _colsum(dh1, 1024, B, 1, 1024)[0]
        dW1 = _wgrad(dh1, 1024, st.g, 1024, B, 1024, 1024)[0]
        dg = _f(B, 1024, dev=dev)
        _dgrad(dh1, 1024, w1.t().contiguous(), 1024, dg, 1024, B, 1024, 1024)
        fg, _ = _feat_backward(st, dg, None, False)
        out = []
        for name in REFINER_PARAMS:
            if name.startswith("feat."):
                dW, db = fg[name[5:]]
                out += [dW.view_as(_get(net, name).weight), db]
            else:
                layer, br = int(name[4]), "rt".index(name[6])
                if layer == 1:
                    out += [dW1[512 * br:512 * (br + 1)].contiguous(), db1[512 * br:512 * (br + 1)].contiguous()]
                elif layer == 2:
                    out += [dW2[br].contiguous(), db2[128 * br:128 * (br + 1)].contiguous()]
                else:
                    dW, db = (g3r, g3t)[br]
                    out += [dW, db]
        return (None, None, None, None, *out)


def posenet_head_train(net, feat, x, choose, obj):
    return PoseNetHeadFn.apply(feat, x, choose, obj, net, *param_list(net, HEAD_PARAMS))


def refiner_train(net, x, emb_pm, obj):
    return RefinerFn.apply(x, emb_pm, obj, net, *param_list(net, REFINER_PARAMS))
