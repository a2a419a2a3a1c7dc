"""Host-side executor of the dense-fusion head (K1) and the refiner (K2) on top of the C ABI.

Algebra (all proven equal to the reference within fp32 rounding):
  * the 1024-wide global feature is never broadcast to the points: its product with columns 384:1408 of
    conv1_{r,t,c} is folded into a PER-CROP BIAS (one small GEMM per crop batch);
  * the (points x 1024) conv6 activation is never stored: the GEMM epilogue pools it;
  * conv4_{r,t,c} (conv3_{r,t} of the refiner) are evaluated for the selected object only;
  * r / t / c towers run as grouped (block-diagonal) GEMMs; concatenations are column slices of one buffer.
Activations are point-major (rows = crops*points, channels contiguous) so every layer is a K-major GEMM.

Data layout in HBM for a chunk of Bc crops (rows = Bc*N): pf (rows,384) = [x1|e1|x2|e2]; h5 (rows,512);
pool partials (Bc,tiles,1024); g (Bc,1024); gbias (Bc,1920); h1 (rows,1920); h2 (rows,768); h3 (rows,384).
A chunk's activations (596 MB for tower 1 alone at the default 128 crops) do NOT stay L2-resident: fewer, larger launches won over
L2 residency when this was measured (16 crops 21.7 ms per step, 128 crops 20.05)."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from ._C import check, lib, ptr, stream


def _w2d(w: torch.Tensor) -> torch.Tensor:
    return w.detach().reshape(w.shape[0], -1).float()


class PackedPoseNetHead:
    """GEMM-ready views / copies of PoseNet's head parameters (lib/network.py:42-49, :77-91)."""

    def __init__(self, net):
        f = net.feat
        self.num_obj = net.num_obj
        self.w_x1, self.b_x1 = _w2d(f.conv1.weight).contiguous(), f.conv1.bias.detach().float().contiguous()
        self.w_e1, self.b_e1 = _w2d(f.e_conv1.weight).contiguous(), f.e_conv1.bias.detach().float().contiguous()
        self.w_x2, self.b_x2 = ops.SplitWeight(_w2d(f.conv2.weight)), f.conv2.bias.detach().float().contiguous()
        self.w_e2, self.b_e2 = ops.SplitWeight(_w2d(f.e_conv2.weight)), f.e_conv2.bias.detach().float().contiguous()
        self.w5, self.b5 = ops.SplitWeight(_w2d(f.conv5.weight)), f.conv5.bias.detach().float().contiguous()
        self.w6, self.b6 = ops.SplitWeight(_w2d(f.conv6.weight)), f.conv6.bias.detach().float().contiguous()
        t1 = [_w2d(getattr(net, f"conv1_{b}").weight) for b in "rtc"]
        self.w1_local = ops.SplitWeight(torch.cat([w[:, :384] for w in t1], 0))       # (1920,384)
        self.w1_global = ops.SplitWeight(torch.cat([w[:, 384:] for w in t1], 0))     # (1920,1024)
        self.b1 = torch.cat([getattr(net, f"conv1_{b}").bias.detach().float() for b in "rtc"]).contiguous()
        self.w2 = ops.SplitWeight(torch.stack([_w2d(getattr(net, f"conv2_{b}").weight) for b in "rtc"]))   # (3,256,640)
        self.b2 = torch.cat([getattr(net, f"conv2_{b}").bias.detach().float() for b in "rtc"]).contiguous()
        self.w3 = ops.SplitWeight(torch.stack([_w2d(getattr(net, f"conv3_{b}").weight) for b in "rtc"]))   # (3,128,256)
        self.b3 = torch.cat([getattr(net, f"conv3_{b}").bias.detach().float() for b in "rtc"]).contiguous()
        self.w4 = [_w2d(getattr(net, f"conv4_{b}").weight).contiguous() for b in "rtc"]
        self.b4 = [getattr(net, f"conv4_{b}").bias.detach().float().contiguous() for b in "rtc"]


class PackedRefiner:
    """GEMM-ready parameters of PoseRefineNet (lib/network.py:139-146, :176-183)."""

    def __init__(self, net):
        f = net.feat
        self.num_obj = net.num_obj
        self.w_x1, self.b_x1 = _w2d(f.conv1.weight).contiguous(), f.conv1.bias.detach().float().contiguous()
        self.w_e1, self.b_e1 = _w2d(f.e_conv1.weight).contiguous(), f.e_conv1.bias.detach().float().contiguous()
        self.w_x2, self.b_x2 = ops.SplitWeight(_w2d(f.conv2.weight)), f.conv2.bias.detach().float().contiguous()
        self.w_e2, self.b_e2 = ops.SplitWeight(_w2d(f.e_conv2.weight)), f.e_conv2.bias.detach().float().contiguous()
        self.w5, self.b5 = ops.SplitWeight(_w2d(f.conv5.weight)), f.conv5.bias.detach().float().contiguous()
        self.w6, self.b6 = ops.SplitWeight(_w2d(f.conv6.weight)), f.conv6.bias.detach().float().contiguous()
        self.w1 = ops.SplitWeight(torch.cat([_w2d(net.conv1_r.weight), _w2d(net.conv1_t.weight)], 0))     # (1024,1024)
        self.b1 = torch.cat([net.conv1_r.bias.detach().float(), net.conv1_t.bias.detach().float()]).contiguous()
        self.w2 = ops.SplitWeight(torch.stack([_w2d(net.conv2_r.weight), _w2d(net.conv2_t.weight)]))       # (2,128,512)
        self.b2 = torch.cat([net.conv2_r.bias.detach().float(), net.conv2_t.bias.detach().float()]).contiguous()
        self.w3 = [_w2d(net.conv3_r.weight).contiguous(), _w2d(net.conv3_t.weight).contiguous()]
        self.b3 = [net.conv3_r.bias.detach().float().contiguous(), net.conv3_t.bias.detach().float().contiguous()]


def param_version(module) -> tuple:
    return tuple((p.data_ptr(), p._version) for p in module.parameters())


class Workspace:
    """Activation scratch for up to `crops` crops of `n` points (allocated once, reused)."""

    def __init__(self, crops: int, n: int, device, towers: bool):
        rows = crops * n
        f = dict(device=device, dtype=torch.float32)
        self.crops, self.n = crops, n
        self.tiles = (n + ops.pool_tile_rows() - 1) // ops.pool_tile_rows()
        self.pf = torch.empty(rows, 384, **f)
        self.h5 = torch.empty(rows, 512, **f)
        self.partial = torch.empty(crops, self.tiles, 1024, **f)
        self.g = torch.empty(crops, 1024, **f)
        if towers:
            self.gbias = torch.empty(crops, 1920, **f)
            self.h1 = torch.empty(rows, 1920, **f)
            self.h2 = torch.empty(rows, 768, **f)
            self.h3 = torch.empty(rows, 384, **f)
        else:
            self.h1 = torch.empty(crops, 1024, **f)
            self.h2 = torch.empty(crops, 256, **f)


def _point_features(w, pf, x, emb_pm, rows, precision, do_emb=True):
    """pf = [relu(conv1 x) | relu(e_conv1 emb) | relu(conv2 .) | relu(e_conv2 .)]  (lib/network.py:54-60)."""
    check(lib.df_xyz_conv(ptr(x), ptr(w.w_x1), ptr(w.b_x1), ptr(pf), 384, rows, stream()), "df_xyz_conv")
    if do_emb:
        ops.gemm(emb_pm, w.w_e1, w.b_e1, pf[:, 64:], M=rows, N=64, K=32, lda=32, ldw=32, ldc=384, relu=True,
                 precision=precision)
    ops.gemm(pf, w.w_x2, w.b_x2, pf[:, 128:], M=rows, N=128, K=64, lda=384, ldw=64, ldc=384, relu=True,
             precision=precision)
    if do_emb:
        ops.gemm(pf[:, 64:], w.w_e2, w.b_e2, pf[:, 256:], M=rows, N=128, K=64, lda=384, ldw=64, ldc=384, relu=True,
                 precision=precision)


# The head and the refiner are written in three phases so that a caller working through many chunks of crops can run
# the per-crop (M = crops) GEMMs -- the folded global-feature bias and the refiner's MLP towers -- ONCE for all chunks:
# at M = 32 they are weight-bandwidth bound launches (4 MB of weights per 2 MFLOP/row), 7% of the pose step when issued
# per chunk.
def head_features_chunk(w: PackedPoseNetHead, ws: Workspace, pf, x, emb_pm, crops, n, g_out, precision="fp32"):
    """Phase A: point features into pf (crops*n,384), pooled global feature into g_out (crops,1024)."""
    rows = crops * n
    _point_features(w, pf, x, emb_pm, rows, precision)
    ops.gemm(pf[:, 128:], w.w5, w.b5, ws.h5, M=rows, N=512, K=256, lda=384, ldw=256, ldc=512, relu=True,
             precision=precision)
    ops.gemm(ws.h5, w.w6, w.b6, None, M=rows, N=1024, K=512, lda=512, ldw=512, ldc=0, relu=True,
             precision=precision, rows_per_crop=n, pool_partial=ws.partial)
    check(lib.df_pool_finish(ptr(ws.partial), ptr(g_out), crops, ws.tiles, 1024, n, stream()), "df_pool_finish")


def head_global_bias(w: PackedPoseNetHead, g, gbias, crops, precision="fp32"):
    """Phase B: folded global feature -> per-crop bias of the first tower layer, (crops,1024) -> (crops,1920)."""
    ops.gemm(g, w.w1_global, w.b1, gbias, M=crops, N=1920, K=1024, lda=1024, ldw=1024, ldc=1920, relu=False,
             precision=precision)


def head_towers_chunk(w: PackedPoseNetHead, ws: Workspace, pf, gbias, obj, crops, n, out_r, out_t, out_c,
                      precision="fp32"):
    """Phase C: r / t / c towers on pf with the per-crop bias, selected object's outputs."""
    rows = crops * n
    ops.gemm(pf, w.w1_local, gbias, ws.h1, M=rows, N=1920, K=384, lda=384, ldw=384, ldc=1920, relu=True,
             precision=precision, bias_crop_stride=1920, rows_per_crop=n)
    ops.gemm(ws.h1, w.w2, w.b2, ws.h2, M=rows, N=256, K=640, lda=1920, ldw=640, ldc=768, relu=True,
             precision=precision, groups=3, a_gs=640, w_gs=256 * 640, bias_gs=256, c_gs=256)
    ops.gemm(ws.h2, w.w3, w.b3, ws.h3, M=rows, N=128, K=256, lda=768, ldw=256, ldc=384, relu=True,
             precision=precision, groups=3, a_gs=256, w_gs=128 * 256, bias_gs=128, c_gs=128)
    check(lib.df_select_out(ptr(ws.h3), 384, ptr(w.w4[0]), ptr(w.b4[0]), ptr(w.w4[1]), ptr(w.b4[1]),
                            ptr(w.w4[2]), ptr(w.b4[2]), ptr(obj), n, w.num_obj, rows, ptr(out_r), ptr(out_t),
                            ptr(out_c), stream()), "df_select_out")


def posenet_head_chunk(w: PackedPoseNetHead, ws: Workspace, x, emb_pm, obj, crops, n, out_r, out_t, out_c,
                       precision="fp32"):
    """x (crops*n,3), emb_pm (crops*n,32), obj (crops,) -> out_r (crops*n,4), out_t (.,3), out_c (.)."""
    head_features_chunk(w, ws, ws.pf, x, emb_pm, crops, n, ws.g, precision)
    head_global_bias(w, ws.g, ws.gbias, crops, precision)
    head_towers_chunk(w, ws, ws.pf, ws.gbias, obj, crops, n, out_r, out_t, out_c, precision)


def refiner_features_chunk(w: PackedRefiner, ws: Workspace, pf, x, emb_pm, crops, n, g_out, precision="fp32",
                           emb_ready=False):
    """x (crops*n,3) re-expressed cloud, emb_pm (crops*n,32) -> pooled feature g_out (crops,1024).
    emb_ready=True: columns e1|e2 of pf already hold this crop batch's embedding features (they do not depend on
    the iteration), only the xyz branch is recomputed."""
    rows = crops * n
    _point_features(w, pf, x, emb_pm, rows, precision, do_emb=not emb_ready)
    ops.gemm(pf, w.w5, w.b5, ws.h5, M=rows, N=512, K=384, lda=384, ldw=384, ldc=512, relu=True, precision=precision)
    ops.gemm(ws.h5, w.w6, w.b6, None, M=rows, N=1024, K=512, lda=512, ldw=512, ldc=0, relu=True,
             precision=precision, rows_per_crop=n, pool_partial=ws.partial)
    check(lib.df_pool_finish(ptr(ws.partial), ptr(g_out), crops, ws.tiles, 1024, n, stream()), "df_pool_finish")


def refiner_mlp(w: PackedRefiner, g, h1, h2, obj, crops, out_r, out_t, precision="fp32"):
    """The two MLP towers on the pooled features: g (crops,1024) -> out_r (crops,4), out_t (crops,3).
    The 1024x1024 layer runs on the tensor cores at any row count; the small second layer stays on the exact-fp32 kernel below
    256 rows (ops.tc_eligible)."""
    ops.gemm(g, w.w1, w.b1, h1, M=crops, N=1024, K=1024, lda=1024, ldw=1024, ldc=1024, relu=True, precision=precision)
    ops.gemm(h1, w.w2, w.b2, h2, M=crops, N=128, K=512, lda=1024, ldw=512, ldc=256, relu=True, precision=precision,
             groups=2, a_gs=512, w_gs=128 * 512, bias_gs=128, c_gs=128)
    check(lib.df_select_out(ptr(h2), 256, ptr(w.w3[0]), ptr(w.b3[0]), ptr(w.w3[1]), ptr(w.b3[1]), None, None,
                            ptr(obj), 1, w.num_obj, crops, ptr(out_r), ptr(out_t), None, stream()), "df_select_out")


def refiner_chunk(w: PackedRefiner, ws: Workspace, x, emb_pm, obj, crops, n, out_r, out_t, precision="fp32",
                  emb_ready=False):
    refiner_features_chunk(w, ws, ws.pf, x, emb_pm, crops, n, ws.g, precision, emb_ready)
    refiner_mlp(w, ws.g, ws.h1, ws.h2, obj, crops, out_r, out_t, precision)
