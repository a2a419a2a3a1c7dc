"""Tensor-level wrappers over the C ABI (one function per entry point of include/densefusion_b200.h).

Each wrapper validates devices / dtypes / shapes, allocates the outputs the C ABI expects the caller to
own, launches on torch's current stream and raises DFError on a non-zero status."""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from . import _C
from ._C import check, f32c, i64c, lib, need_cuda, ptr, stream

# "hybrid": TF32 main term + the two ~2^-11 correction terms in bf16 (8 instead of 12 MMAs per k-block), fp32-parity like
# "3xtf32"; generation-2 tensor-core kernels only
# "hybrid16": fp16 main term + bf16 correction terms (6 MMAs per k-block), weight planes packed once (df_pack_f16_pairs).
# (code 5 was "hybrid16w", the same arithmetic with the fp32 weight tile split on chip: bit-identical, measured slower -- tower-1
# 0.41 vs 0.30 ms, profiles/r2_c3_gemm_ab.jsonl -- and removed again.)
# "hybrid16s": every term on fp16 operands, two planes per operand (x = fp16(x s) + fp16(x s - fp16(x s)), power-of-two scales s):
# 4 instead of 6 bytes per weight element through the SM's fabric port and half the TMEM per A stage.  The weight scale comes from
# the tensor's own maximum (df_pack_f16s); the activation scale is derived inside the kernel from a 4096-element sample of the operand
# (gemm_tc.cu, "activation scale"), or fixed per call with `a_log2` (tests).
PRECISIONS = {"fp32": 0, "3xtf32": 1, "tf32": 2, "hybrid": 3, "hybrid16": 4, "hybrid16s": 6}


def _prec_code(mode: int, short_runs: bool, a_log2=None) -> int:
    """`precision` argument of df_gemm_tc / df_conv_tc: mode | run-length flag | hybrid16s activation scale (bits 16..23: 0 = sampled
    by the kernel, otherwise the signed log2 of a fixed scale, -128 standing for 2^0)."""
    code = mode | (SHORT_RUNS if short_runs else 0)
    if mode == 6 and a_log2 is not None:
        code |= ((int(a_log2) & 0xff) if a_log2 else 0x80) << 16
    return code


def sym_mask(sym_list: Iterable[int]) -> int:
    m = 0
    for s in sym_list:
        s = int(s)
        if s < 0 or s >= 64:
            raise _C.DFError("object ids in sym_list must be in [0, 64)")
        m |= 1 << s
    return m


# ---- K4 -----------------------------------------------------------------------------------------
def knn(ref: torch.Tensor, query: torch.Tensor, k: int = 1, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ref (B,D,R), query (B,D,Q) -> int64 (B,k,Q), 1-based (lib/knn/__init__.py:15-23)."""
    need_cuda(ref, query)
    if ref.dim() != 3 or query.dim() != 3:
        raise _C.DFError("ref_tensor / query_tensor: 3D Tensor expected")
    if ref.shape[0] != query.shape[0] or ref.shape[1] != query.shape[1]:
        raise _C.DFError("input sizes must match")
    ref, query = f32c(ref), f32c(query)
    B, D, R = ref.shape
    Q = query.shape[2]
    if out is None:
        out = torch.empty(B, k, Q, dtype=torch.int64, device=ref.device)
    elif out.dim() != 3 or out.shape[2] != Q or out.shape[0] != B or out.dtype != torch.int64 or not out.is_contiguous():
        raise _C.DFError("idx_tensor: contiguous int64 (B,k,Q) expected")
    check(lib.df_knn(ptr(ref), ptr(query), ptr(out), B, D, R, Q, out.shape[1], stream()), "df_knn")
    return out


# ---- K3 -----------------------------------------------------------------------------------------
class LossState:
    """Outputs + saved-for-backward buffers of one df_loss_forward call."""
    __slots__ = ("loss", "dis_sel", "which", "new_points", "new_target", "dis_all", "sum_u", "sum_um",
                 "dbg_pred", "dbg_nn")


_TICKETS = {}


def _tickets(device, B):
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    t = _TICKETS.get(key)
    if t is None or t.numel() < B:
        t = torch.zeros(max(B, 1024), dtype=torch.int32, device=device)
        _TICKETS[key] = t
    return t


def loss_forward(pred_r, pred_t, pred_c, target, model_points, hyp_points, points, idx, sym: int,
                 allow_sym: bool, w: float, debug: bool = False) -> LossState:
    """Batched fused loss.  pred_r (B,P,4), pred_t (B,P,3), pred_c (B,P,1)|None, target/model_points (B,M,3),
    hyp_points (B,P,3)|None, points (B,N,3), idx (B,)|(B,1)."""
    need_cuda(pred_r, pred_t, pred_c, target, model_points, hyp_points, points, idx)
    pred_r, pred_t, target, model_points, points = map(f32c, (pred_r, pred_t, target, model_points, points))
    pred_c = None if pred_c is None else f32c(pred_c)
    hyp_points = None if hyp_points is None else f32c(hyp_points)
    idx = i64c(idx).view(-1)
    B, P = pred_r.shape[0], pred_r.shape[1]
    M, N = target.shape[1], points.shape[1]
    if model_points.shape[1] != M or pred_t.shape[:2] != (B, P) or idx.numel() != B or points.shape[0] != B:
        raise _C.DFError("loss: inconsistent shapes")
    dev, f = pred_r.device, torch.float32
    st = LossState()
    st.dis_all = torch.empty(B, P, device=dev, dtype=f)
    st.sum_u = torch.empty(B, P, 3, device=dev, dtype=f)
    st.sum_um = torch.empty(B, P, 9, device=dev, dtype=f)
    st.loss = torch.empty(B, device=dev, dtype=f)
    st.dis_sel = torch.empty(B, device=dev, dtype=f)
    st.which = torch.empty(B, device=dev, dtype=torch.int64)
    st.new_points = torch.empty(B, N, 3, device=dev, dtype=f)
    st.new_target = torch.empty(B, M, 3, device=dev, dtype=f)
    st.dbg_pred = torch.empty(B, P, M, 3, device=dev, dtype=f) if debug else None
    st.dbg_nn = torch.empty(B, P, M, device=dev, dtype=torch.int32) if debug else None
    check(lib.df_loss_forward(ptr(pred_r), ptr(pred_t), ptr(pred_c), ptr(target), ptr(model_points),
                              ptr(hyp_points), ptr(points), ptr(idx), sym, 1 if allow_sym else 0, float(w),
                              B, P, M, N, ptr(st.dis_all), ptr(st.sum_u), ptr(st.sum_um), ptr(st.loss),
                              ptr(st.dis_sel), ptr(st.which), ptr(st.new_points), ptr(st.new_target),
                              ptr(_tickets(dev, B)), ptr(st.dbg_pred), ptr(st.dbg_nn), stream()),
          "df_loss_forward")
    return st


def loss_backward(pred_r, pred_c, st: LossState, g_loss, g_dis, w: float):
    pred_r = f32c(pred_r)
    pred_c = None if pred_c is None else f32c(pred_c)
    B, P = pred_r.shape[0], pred_r.shape[1]
    g_r = torch.empty(B, P, 4, device=pred_r.device, dtype=torch.float32)
    g_t = torch.empty(B, P, 3, device=pred_r.device, dtype=torch.float32)
    g_c = torch.empty(B, P, 1, device=pred_r.device, dtype=torch.float32) if pred_c is not None else None
    g_loss = None if g_loss is None else f32c(g_loss).view(-1)
    g_dis = None if g_dis is None else f32c(g_dis).view(-1)
    check(lib.df_loss_backward(ptr(pred_r), ptr(pred_c), ptr(st.dis_all), ptr(st.sum_u), ptr(st.sum_um),
                               ptr(st.which), ptr(g_loss), ptr(g_dis), float(w), B, P, ptr(g_r), ptr(g_t),
                               ptr(g_c), stream()), "df_loss_backward")
    return g_r, g_t, g_c


# ---- K1 / K2 building blocks ------------------------------------------------------------------
class SplitWeight:
    """A torch (N,K) [or stacked (G,N,K)] weight with its TF32 hi/lo halves for the tensor-core path."""
    __slots__ = ("w", "hi", "lo", "bf", "h16", "s16")

    def __init__(self, w: torch.Tensor):
        self.w = f32c(w.detach())
        self.hi = self.lo = self.bf = self.h16 = self.s16 = None

    def operands(self, mode: int):
        """The two weight operands df_gemm_tc / df_conv_tc expect for a PRECISIONS code (1 3xtf32, 2 tf32, 3 hybrid, 4 hybrid16,
        6 hybrid16s: the packed planes and their scale record)."""
        if mode == 6:
            return self.planes16s()
        return self.pairs16() if mode == 4 else (self.pairs() if mode == 3 else self.split())

    def pairs16(self):
        """The two packed 16-bit tensors of the hybrid16 mode ([fp16(w) | bf16(w)] per k-block, and bf16(w - fp16(w)) row-major)."""
        if self.h16 is None:
            need_cuda(self.w)
            K = self.w.shape[-1]
            rows = self.w.numel() // K
            t1 = torch.empty(rows, K, device=self.w.device, dtype=torch.float32)
            t2 = torch.empty(rows, K // 2, device=self.w.device, dtype=torch.float32)        # K bf16 per row
            check(lib.df_pack_f16_pairs(ptr(self.w), ptr(t1), ptr(t2), rows, K, stream()), "df_pack_f16_pairs")
            self.h16 = (t1, t2)
        return self.h16

    def planes16s(self):
        """hybrid16s: the packed planes ([fp16(w s) x32 | fp16(w s - fp16(w s)) x32] per row and k-block) and the 4-float scale record
        (1/s, s, scratch) made by df_pack_f16s on the device (no host round trip: usable under graph capture)."""
        if self.s16 is None:
            need_cuda(self.w)
            K = self.w.shape[-1]
            rows = self.w.numel() // K
            planes = torch.empty(rows, K, device=self.w.device, dtype=torch.float32)
            scale = torch.empty(4, device=self.w.device, dtype=torch.float32)
            check(lib.df_pack_f16s(ptr(self.w), ptr(planes), ptr(scale), rows, K, stream()), "df_pack_f16s")
            self.s16 = (planes, scale)
        return self.s16

    def pairs(self):
        """hi (TF32-exact fp32) and the packed bf16 pair tensor of the hybrid mode (same byte size as the weight)."""
        hi, _ = self.split()
        if self.bf is None:
            K = self.w.shape[-1]
            rows = self.w.numel() // K
            self.bf = torch.empty(rows, K, device=self.w.device, dtype=torch.float32)     # 2K bf16 per row
            check(lib.df_pack_bf16_pairs(ptr(self.w), ptr(self.bf), rows, K, stream()), "df_pack_bf16_pairs")
        return hi, self.bf

    def split(self):
        if self.hi is None:
            need_cuda(self.w)
            self.hi, self.lo = torch.empty_like(self.w), torch.empty_like(self.w)
            check(lib.df_split_tf32(ptr(self.w), ptr(self.hi), ptr(self.lo), self.w.numel(), stream()), "df_split_tf32")
        return self.hi, self.lo


TC_VARIANT = 0          # 0 auto; 1/2/3 force a kernel variant (bring-up / tests)


def tc_eligible(M: int, N: int, K: int) -> bool:
    """Shapes worth a tensor-core launch; the rest stays on the exact-fp32 kernel in every mode.  Short operands (M < 256: the
    refiner's towers, the pooled pyramid branches) are one partially filled tile per CTA pair -- still 2-5x faster than the
    FFMA kernel once the weight matrix is large (measured: M=96 N=1024 K=512 15 vs 64 us; M=32 N=512 K=128 a tie)."""
    return (M >= 256 or N * K >= 512 * 512) and K % 32 == 0 and K >= 64 and N % 64 == 0


SHORT_RUNS = 9 << 8     # precision flag of df_gemm_tc / df_conv_tc: accumulation runs of 108 MMA instructions (training path)


def gemm(A, W, bias, C, *, M, N, K, lda, ldw, ldc, relu, precision="fp32", bias_crop_stride=0, rows_per_crop=0,
         groups=1, a_gs=0, w_gs=0, bias_gs=0, c_gs=0, pool_partial=None, short_runs=False, a_log2=None):
    """Raw strided GEMM launch on pre-allocated buffers (see df_gemm_fp32 / df_gemm_tc in the header).
    W: tensor or SplitWeight."""
    mode = PRECISIONS[precision]
    sw = W if isinstance(W, SplitWeight) else None
    if mode != 0 and sw is not None and tc_eligible(M, N, K) and (groups == 1 or w_gs == N * ldw):
        hi, lo = sw.operands(mode)
        st = lib.df_gemm_tc(ptr(A), lda, ptr(hi), ptr(lo), ldw, ptr(bias), bias_crop_stride, ptr(C), ldc, M, N, K,
                            1 if relu else 0, rows_per_crop, groups, a_gs, bias_gs, c_gs, ptr(pool_partial),
                            _prec_code(mode, short_runs, a_log2), TC_VARIANT, stream())
        if st != -2:                 # -2: a shape the TMA boxes cannot address -> the exact-fp32 kernel below, like tc_eligible
            check(st, "df_gemm_tc")
            return
    Wt = sw.w if sw is not None else W
    st = lib.df_gemm_fp32(ptr(A), lda, ptr(Wt), ldw, ptr(bias), bias_crop_stride, ptr(C), ldc, M, N, K,
                          1 if relu else 0, rows_per_crop, groups, a_gs, w_gs, bias_gs, c_gs,
                          ptr(pool_partial), stream())
    check(st, "df_gemm_fp32")


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], relu: bool = False,
           precision: str = "fp32") -> torch.Tensor:
    """y = act(x @ weight.T + bias) for 2-D row-major x (rows, K) and torch (N, K) weight."""
    need_cuda(x, weight, bias)
    x, weight = f32c(x), f32c(weight)
    M, K = x.shape
    N = weight.shape[0]
    y = torch.empty(M, N, device=x.device, dtype=torch.float32)
    if precision != "fp32":
        weight = SplitWeight(weight)
    gemm(x, weight, None if bias is None else f32c(bias), y, M=M, N=N, K=K, lda=K, ldw=K, ldc=N, relu=relu,
         precision=precision)
    return y


def pool_tile_rows() -> int:
    return int(lib.df_gemm_rows_per_pool_tile())


def gather_embedding(feat: torch.Tensor, choose: torch.Tensor, want_pm=True, want_cm=True):
    """feat (B,32,H,W) any strides (NCHW or channels_last), choose (B,1,N)|(B,N) i64."""
    need_cuda(feat, choose)
    if feat.dtype != torch.float32:
        feat = feat.float()
    B, C, H, W = feat.shape
    if C != 32:
        raise _C.DFError("gather_embedding: 32 channels expected")
    sb, sc, sh, sw = feat.stride()
    if sh != W * sw:                       # pixels must be addressable with one stride
        feat = feat.contiguous()
        sb, sc, sh, sw = feat.stride()
    choose = i64c(choose).view(B, -1)
    N = choose.shape[1]
    pm = torch.empty(B * N, 32, device=feat.device, dtype=torch.float32) if want_pm else None
    cm = torch.empty(B, 32, N, device=feat.device, dtype=torch.float32) if want_cm else None
    check(lib.df_gather_embedding(ptr(feat), ptr(choose), ptr(pm), ptr(cm), sb, sc, sw, B, N, H * W, stream()),
          "df_gather_embedding")
    return pm, cm


# ---- K5 -----------------------------------------------------------------------------------------
def select_pose(pred_r, pred_t, pred_c, points):
    need_cuda(pred_r, pred_t, pred_c, points)
    pred_r, pred_t, pred_c, points = map(f32c, (pred_r, pred_t, pred_c, points))
    B, N = pred_r.shape[0], pred_r.shape[1]
    pose = torch.empty(B, 7, device=pred_r.device, dtype=torch.float64)
    which = torch.empty(B, device=pred_r.device, dtype=torch.int64)
    check(lib.df_select_pose(ptr(pred_r), ptr(pred_t), ptr(pred_c), ptr(points), B, N, ptr(pose), ptr(which),
                             stream()), "df_select_pose")
    return pose, which


def cloud_transform(cloud, pose, out=None):
    need_cuda(cloud, pose)
    cloud = f32c(cloud)
    B, N = cloud.shape[0], cloud.shape[1]
    if pose.dtype != torch.float64 or not pose.is_contiguous() or pose.shape != (B, 7):
        raise _C.DFError("pose: contiguous float64 (B,7) expected")
    if out is None:
        out = torch.empty_like(cloud)
    check(lib.df_cloud_transform(ptr(cloud), ptr(pose), ptr(out), B, N, stream()), "df_cloud_transform")
    return out


def pose_compose_(pose, r2, t2):
    need_cuda(pose, r2, t2)
    r2, t2 = f32c(r2), f32c(t2)
    B = pose.shape[0]
    if pose.dtype != torch.float64 or not pose.is_contiguous() or pose.shape != (B, 7):
        raise _C.DFError("pose: contiguous float64 (B,7) expected")
    check(lib.df_pose_compose(ptr(pose), ptr(r2), ptr(t2), B, stream()), "df_pose_compose")
    return pose


# ---- encoder helper -----------------------------------------------------------------------------------
def upsample_bilinear(x: torch.Tensor, size, align_corners: bool) -> torch.Tensor:
    """NCHW fp32 CUDA tensor -> (N,C,size[0],size[1]); same arithmetic as F.interpolate(mode='bilinear')."""
    need_cuda(x)
    x = f32c(x)
    n, c, h, w = x.shape
    out = torch.empty(n, c, int(size[0]), int(size[1]), device=x.device, dtype=torch.float32)
    check(lib.df_upsample_bilinear(ptr(x), ptr(out), n * c, h, w, int(size[0]), int(size[1]), 1 if align_corners else 0,
                                   stream()), "df_upsample_bilinear")
    return out
