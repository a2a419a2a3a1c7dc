"""Drop-in `Loss_refine` (reference: lib/loss_refiner.py:12-74) on the fused K3 kernel (single-hypothesis
form: pred = model . R^T + t, symmetric objects always matched through the 1-NN, no confidence term).

Returns (dis (1,), new_points (1,N,3), new_target (1,M,3)); differentiable w.r.t. pred_r / pred_t.
Additive: B crops at once ((B,4), (B,3)) -> dis (B,)."""
from __future__ import annotations

import torch
from torch.nn.modules.loss import _Loss

from .. import ops


class _FusedRefineLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_r, pred_t, target, model_points, idx, points, sym):
        B = pred_r.shape[0]
        st = ops.loss_forward(pred_r.detach().view(B, 1, 4), pred_t.detach().view(B, 1, 3), None, target,
                              model_points, None, points, idx, sym, True, 0.0)
        ctx.st = st
        ctx.save_for_backward(pred_r.detach())
        ctx.mark_non_differentiable(st.new_points, st.new_target)
        return st.dis_sel, st.new_points, st.new_target

    @staticmethod
    def backward(ctx, g_dis, _gp, _gt):
        (pred_r,) = ctx.saved_tensors
        B = pred_r.shape[0]
        g_r, g_t, _ = ops.loss_backward(pred_r.view(B, 1, 4), None, ctx.st, None, g_dis, 0.0)
        return g_r.view_as(pred_r), g_t.view(B, 3), None, None, None, None, None


def loss_calculation(pred_r, pred_t, target, model_points, idx, points, num_point_mesh, sym_list):
    if target.shape[1] != num_point_mesh or model_points.shape[1] != num_point_mesh:
        raise RuntimeError("target / model_points must hold num_points_mesh points")
    B = target.shape[0]
    dis, new_points, new_target = _FusedRefineLoss.apply(pred_r.reshape(B, 4), pred_t.reshape(B, 3), target,
                                                         model_points, idx, points, ops.sym_mask(sym_list))
    return dis, new_points, new_target


class Loss_refine(_Loss):
    def __init__(self, num_points_mesh, sym_list):
        super().__init__()
        self.num_pt_mesh = num_points_mesh
        self.sym_list = sym_list

    def forward(self, pred_r, pred_t, target, model_points, idx, points):
        return loss_calculation(pred_r, pred_t, target, model_points, idx, points, self.num_pt_mesh, self.sym_list)
