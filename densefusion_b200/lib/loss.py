"""Drop-in `Loss` (reference: lib/loss.py:13-82) on the fused K3 kernel.

Same constructor and forward signature; returns (loss, dis, new_points, new_target) with the reference's
shapes for bs == 1 -- loss and dis are 0-d tensors, new_points (1,N,3), new_target (1,M,3), the last two
detached.  Differentiable w.r.t. pred_r / pred_t / pred_c through df_loss_backward.  Unlike the reference
there is no host synchronisation (`idx[0].item()`, lib/loss.py:41): the symmetric branch is chosen on the
device from `idx` and a bit mask of `sym_list`.

Additive: a batch of B crops is accepted ((B,N,4) ...); then loss / dis have shape (B,)."""
from __future__ import annotations

import torch
from torch.nn.modules.loss import _Loss

from .. import ops


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_r, pred_t, pred_c, target, model_points, idx, points, w, allow_sym, sym):
        st = ops.loss_forward(pred_r.detach(), pred_t.detach(), pred_c.detach(), target, model_points, points,
                              points, idx, sym, allow_sym, w)
        ctx.st, ctx.w = st, w
        ctx.save_for_backward(pred_r.detach(), pred_c.detach())
        ctx.mark_non_differentiable(st.new_points, st.new_target, st.which)
        return st.loss, st.dis_sel, st.new_points, st.new_target, st.which

    @staticmethod
    def backward(ctx, g_loss, g_dis, _gp, _gt, _gw):
        pred_r, pred_c = ctx.saved_tensors
        g_r, g_t, g_c = ops.loss_backward(pred_r, pred_c, ctx.st, g_loss, g_dis, ctx.w)
        return g_r.view_as(pred_r), g_t.view(pred_r.shape[0], pred_r.shape[1], 3), g_c, None, None, None, None, None, None, None


def loss_calculation(pred_r, pred_t, pred_c, target, model_points, idx, points, w, refine, num_point_mesh, sym_list):
    bs, num_p = pred_c.shape[0], pred_c.shape[1]
    if target.shape[1] != num_point_mesh or model_points.shape[1] != num_point_mesh:
        raise RuntimeError("target / model_points must hold num_points_mesh points")
    loss, dis, new_points, new_target, _ = _FusedLoss.apply(
        pred_r, pred_t, pred_c.reshape(bs, num_p, 1), target, model_points, idx, points, float(w),
        not bool(refine), ops.sym_mask(sym_list))
    if bs == 1:
        return loss[0], dis[0], new_points, new_target
    return loss, dis, new_points, new_target


class Loss(_Loss):
    def __init__(self, num_points_mesh, sym_list):
        super().__init__()
        self.num_pt_mesh = num_points_mesh
        self.sym_list = sym_list

    def forward(self, pred_r, pred_t, pred_c, target, model_points, idx, points, w, refine):
        return loss_calculation(pred_r, pred_t, pred_c, target, model_points, idx, points, w, refine,
                                self.num_pt_mesh, self.sym_list)
