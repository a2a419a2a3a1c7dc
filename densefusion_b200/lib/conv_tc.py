"""Training-mode convolution of the colour encoder on the tensor cores (training side).

`conv2d(module, x)` stands in for `module(x)` inside lib/pspnet.py / lib/extractors.py.  For the stride-1 3x3 / 1x1
convolutions with a multiple of 32 input channels (all but the three stride-2 layers) on CUDA with autograd on, forward
and the data gradient run on df_conv_tc -- the implicit-GEMM tcgen05 kernel in the fp32-parity "hybrid16" arithmetic:
    y  = conv(x, W)                  weights repacked (Cout, taps*Cin), split every call (they change every step)
    dx = conv(dy, rot180(W)^T)       same kernel, same padding / dilation (exact for stride 1)
    dW = dy^T x_shifted              df_conv_wgrad_tc: one 3xTF32 GEMM whose reduction runs over the zero-padded pixel axis
The three stride-2 layers (conv1 7x7/2, layer2.0.conv1 3x3/2 and its 1x1/2 projection) run as explicit patch matrices on the same
GEMM kernel (ConvS2Fn).  Forward arithmetic `PRECISION` ("hybrid16"), data gradients `GRAD_PRECISION` ("hybrid": gradients need
the fp32 exponent range).  ReLU and the residual blocks' skip connections are epilogues of the convolution launch
(`conv2d(m, x, act=1, residual=skip)`; backward masks dy with [y > 0] and hands the masked gradient to the skip input).
Activations are channels_last, i.e. physically the NHWC layout the kernel wants; the remaining encoder ops (pooling, pyramid
concat, PReLU, Dropout2d, log-softmax) are own kernels as well (lib/ew.py)."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import ops
from .._C import check, lib, ptr, stream

ENABLED = True          # module switch (tests compare against the pure torch path)
WGRAD_TC = True         # weight gradients on df_conv_wgrad_tc (3xTF32 GEMM over the pixels); False: aten.convolution_backward
# The three stride-2 layers as explicit patch matrices (ConvS2Fn): on by default since round 2.  (Round 1 left it off because the
# graphed step's warm-up faulted with an illegal address whenever it was on; the fault was not in these kernels: the tower-1 GEMM
# epilogue read its per-crop bias one row past the buffer in the masked tail of the last M tile, and whether that stray read hit
# unmapped memory depended on the allocation pattern -- fixed in gemm_tc.cu, see tests/test_head_gpu.py::
# test_per_crop_bias_is_not_read_past_its_last_row.)  DF_STRIDE2_TC=0 puts the three layers back on aten.convolution.
STRIDE2_TC = os.environ.get("DF_STRIDE2_TC", "1") == "1"
_S2_ONLY = tuple(int(v) for v in os.environ.get("DF_S2_ONLY", "7,3,1").split(",") if v)     # debug: kernel sizes ConvS2Fn takes
_S2_SYNC = os.environ.get("DF_S2_SYNC", "0") == "1"                                          # debug: synchronise after every call


def _s2_sync(what):
    if _S2_SYNC:
        try:
            torch.cuda.synchronize()
        except Exception as e:
            raise RuntimeError(f"ConvS2Fn: CUDA error surfaced after {what}") from e
PRECISION = "hybrid16"
GRAD_PRECISION = "hybrid"   # data-gradient convolutions: their A operand is a gradient (often < 6e-5 in magnitude, below fp16's
                            # normal range), so the main term must keep the fp32 exponent -- see training.py


def _nhwc(x: torch.Tensor) -> torch.Tensor:
    """(B,C,H,W) -> a channels_last tensor whose storage is NHWC-contiguous."""
    return x if x.is_contiguous(memory_format=torch.channels_last) and x.stride(1) == 1 else x.contiguous(memory_format=torch.channels_last)


# Packed operands of the current optimiser step, keyed by (weight storage, rotate, mode).  The weights only change in the
# optimiser, so inside one step (several crop-size buckets, forward + backward) every weight is packed once per form instead
# of once per call.  Off by default: only the trainer, which knows where a step begins, switches it on (`weight_cache`).
_cache = None


class weight_cache:
    """Context manager around ONE optimiser step's forward + backward: packed convolution weights are reused inside it."""

    def __enter__(self):
        global _cache
        self.prev, _cache = _cache, {}
        return self

    def __exit__(self, *exc):
        global _cache
        _cache = self.prev
        return False


def _pack(weight: torch.Tensor, rotate: bool, mode: int):
    """(Cout,Cin,k,k) -> the split GEMM operand of the forward (rotate=False) or data-gradient (rotate=True) convolution,
    one launch (the weights change every optimiser step, so this runs per call unless a `weight_cache` is active)."""
    if _cache is not None:
        key = (weight.data_ptr(), tuple(weight.shape), rotate, mode)
        hit = _cache.get(key)
        if hit is None:
            hit = _cache[key] = _pack_now(weight, rotate, mode)
        return hit
    return _pack_now(weight, rotate, mode)


def _pack_now(weight: torch.Tensor, rotate: bool, mode: int):
    cout, cin, k, _ = weight.shape
    rows, kt = (cin, k * k * cout) if rotate else (cout, k * k * cin)
    w = weight.detach().float().contiguous()
    hi = torch.empty(rows, kt, device=w.device, dtype=torch.float32)
    if mode == 6:                                           # hybrid16s: the repacked fp32 weights -> two fp16 planes + scale record
        check(lib.df_pack_conv_weight(ptr(w), ptr(hi), None, None, cout, cin, k * k, 1 if rotate else 0, stream()),
              "df_pack_conv_weight")
        planes, scale = torch.empty_like(hi), torch.empty(4, device=w.device, dtype=torch.float32)
        check(lib.df_pack_f16s(ptr(hi), ptr(planes), ptr(scale), rows, kt, stream()), "df_pack_f16s")
        return planes, scale
    if mode == 4:                                           # hybrid16: [fp16(w) | bf16(w)] per k-block and bf16(w - fp16(w))
        second = torch.empty(rows, kt // 2, device=w.device, dtype=torch.float32)
        check(lib.df_pack_conv_weight16(ptr(w), ptr(hi), ptr(second), cout, cin, k * k, 1 if rotate else 0, stream()),
              "df_pack_conv_weight16")
        return hi, second
    second = torch.empty(rows, kt, device=w.device, dtype=torch.float32)
    check(lib.df_pack_conv_weight(ptr(w), ptr(hi), ptr(second) if mode != 3 else None, ptr(second) if mode == 3 else None,
                                  cout, cin, k * k, 1 if rotate else 0, stream()), "df_pack_conv_weight")
    return hi, second


def _launch(x, packed, bias, cout, taps, dil, mode, act=0, residual=None):
    """y = act(conv(x) + bias + residual): ReLU (act = 1) and the skip connection are epilogues of the convolution kernel."""
    b, cin, h, w = x.shape
    y = torch.empty(b, cout, h, w, device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
    hi, lo = packed
    check(lib.df_conv_tc(ptr(x), b, h, w, cin, cin, ptr(hi), ptr(lo), taps, dil, ptr(bias), ptr(residual), cout if residual is not None else 0,
                         None, act, ptr(y), cout, cout, mode | ops.SHORT_RUNS, stream()), "df_conv_tc")
    return y


def _wgrad_tc(x, dy, cout, cin, taps, dil):
    """dW (Cout,Cin,k,k) from channels_last x (B,Cin,H,W) and dy (B,Cout,H,W): df_conv_wgrad_tc + the tap-major -> torch reorder."""
    b, _, h, w = x.shape
    n = int(lib.df_conv_wgrad_scratch_floats(b, h, w, cin, cout, taps, dil))
    scratch = torch.empty(n, device=x.device, dtype=torch.float32)
    out = torch.empty(cout, taps * cin, device=x.device, dtype=torch.float32)
    check(lib.df_conv_wgrad_tc(ptr(x), cin, ptr(dy), cout, b, h, w, cin, cout, taps, dil, ptr(scratch), ptr(out), stream()),
          "df_conv_wgrad_tc")
    k = 3 if taps == 9 else 1
    return out.view(cout, taps, cin).permute(0, 2, 1).reshape(cout, cin, k, k)


class ConvTCFn(torch.autograd.Function):
    """y = act(conv(x, weight) + bias + residual), act = 0 | 1 (ReLU): the BasicBlock's `relu(conv2(.) + skip)` (lib/extractors.py:
    57-70) is ONE kernel launch forward; backward masks dy with [y > 0] (own kernel) and hands the masked gradient to the skip input."""

    @staticmethod
    def forward(ctx, x, weight, bias, dilation, act=0, residual=None):
        x = _nhwc(x.detach().float())
        cout, cin, k, _ = weight.shape
        mode = ops.PRECISIONS[PRECISION]
        packed = _pack(weight, False, mode)
        res = None if residual is None else _nhwc(residual.detach().float())
        y = _launch(x, packed, None if bias is None else bias.detach().float().contiguous(), cout, k * k, dilation, mode, act, res)
        if act:
            ctx.save_for_backward(x, weight, y)
        else:
            ctx.save_for_backward(x, weight)
        ctx.dilation, ctx.has_bias, ctx.mode, ctx.act, ctx.has_res = dilation, bias is not None, mode, act, residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import ew
        x, weight = ctx.saved_tensors[:2]
        cout, cin, k, _ = weight.shape
        dy = _nhwc(dy.float())
        if ctx.act:
            dy = ew.relu_mask(dy, ctx.saved_tensors[2])
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # data gradient = convolution of dy with the 180-degree rotated, in/out-transposed kernel
            gmode = ops.PRECISIONS[GRAD_PRECISION] if ctx.mode >= 4 else ctx.mode
            dx = _launch(dy, _pack(weight, True, gmode), None, cin, k * k, ctx.dilation, gmode)
        if ctx.needs_input_grad[1] and WGRAD_TC and cin % 64 == 0:
            dw = _wgrad_tc(x, dy, cout, cin, k * k, ctx.dilation)
        elif ctx.needs_input_grad[1]:
            pad = ctx.dilation * (k // 2)
            dw = torch.ops.aten.convolution_backward(dy, x, weight, None, [1, 1], [pad, pad], [ctx.dilation, ctx.dilation],
                                                     False, [0, 0], 1, [False, True, False])[1]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            from . import ew
            db = ew.colsum(dy)
        dres = dy if (ctx.has_res and ctx.needs_input_grad[5]) else None
        return dx, dw, db, None, None, dres


K_CONV1 = 192           # 3*7*7 = 147 im2col columns, zero-padded to a multiple of 64 (df_conv_wgrad_tc wants Cin % 64 == 0)


class ConvS2Fn(torch.autograd.Function):
    """The three stride-2 convolutions (conv1 7x7/2 on the NCHW image, layer2.0.conv1 3x3/2 and its 1x1/2 projection on
    channels_last activations) as explicit patch matrices: y = A W^T with A = im2col(x) (df_enc_im2col_conv1 /
    df_enc_im2col_s2 / the even pixels), dW = dy^T A (df_conv_wgrad_tc), dx = col2im(dy W) (df_enc_col2im_s2)."""

    @staticmethod
    def _patches(x, k):
        b, cin, h, w = x.shape
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        m = b * ho * wo
        if k == 7:
            x = x.detach().float().contiguous()                             # the 3-channel image, NCHW as the reference feeds it
            a = torch.empty(m, K_CONV1, device=x.device, dtype=torch.float32)
            check(lib.df_enc_im2col_conv1(ptr(x), ptr(a), b, h, w, K_CONV1, stream()), "df_enc_im2col_conv1")
        elif k == 3:
            xn = _nhwc(x.detach().float())
            a = torch.empty(m, 9 * cin, device=x.device, dtype=torch.float32)
            check(lib.df_enc_im2col_s2(ptr(xn), ptr(a), b, h, w, cin, stream()), "df_enc_im2col_s2")
        else:
            a = x.detach().float()[:, :, ::2, ::2].permute(0, 2, 3, 1).reshape(m, cin).contiguous()
        return a, ho, wo

    @staticmethod
    def _matrix(weight):
        cout, cin, k, _ = weight.shape
        w = weight.detach().float()
        if k == 7:
            wm = torch.zeros(cout, K_CONV1, device=w.device, dtype=torch.float32)
            wm[:, :147] = w.reshape(cout, 147)
            return wm
        return w.permute(0, 2, 3, 1).reshape(cout, k * k * cin).contiguous()   # tap-major, channels fastest (as im2col_s2)

    @staticmethod
    def forward(ctx, x, weight, act=0):
        cout, cin, k, _ = weight.shape
        b = x.shape[0]
        a, ho, wo = ConvS2Fn._patches(x, k)
        wm = ConvS2Fn._matrix(weight)
        y = torch.empty(b * ho * wo, cout, device=x.device, dtype=torch.float32)
        ops.gemm(a, ops.SplitWeight(wm), None, y, M=a.shape[0], N=cout, K=a.shape[1], lda=a.shape[1], ldw=a.shape[1], ldc=cout,
                 relu=bool(act), precision=PRECISION, short_runs=True)
        if act:
            ctx.save_for_backward(a, weight, y)
        else:
            ctx.save_for_backward(a, weight)
        ctx.in_shape, ctx.act = tuple(x.shape), act
        _s2_sync(f"forward k={k} x={tuple(x.shape)}")
        return y.view(b, ho, wo, cout).permute(0, 3, 1, 2)                   # channels_last storage, NCHW shape

    @staticmethod
    def backward(ctx, dy):
        a, weight = ctx.saved_tensors[:2]
        cout, cin, k, _ = weight.shape
        b, _, h, w = ctx.in_shape
        m, kk = a.shape
        dy2 = dy.float().permute(0, 2, 3, 1).reshape(m, cout)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        if ctx.act:                                                          # ReLU fused into the GEMM epilogue: mask with [y > 0]
            masked = torch.empty_like(dy2)
            check(lib.df_ew_relu_mask(ptr(dy2), ptr(ctx.saved_tensors[2]), ptr(masked), cout, cout, m, stream()), "df_ew_relu_mask")
            dy2 = masked
        dx = dw = None
        if ctx.needs_input_grad[1]:
            n = int(lib.df_conv_wgrad_scratch_floats(1, 1, m, kk, cout, 1, 1))
            scratch = torch.empty(n, device=a.device, dtype=torch.float32)
            dwm = torch.empty(cout, kk, device=a.device, dtype=torch.float32)
            check(lib.df_conv_wgrad_tc(ptr(a), kk, ptr(dy2), cout, 1, 1, m, kk, cout, 1, 1, ptr(scratch), ptr(dwm), stream()),
                  "df_conv_wgrad_tc")
            dw = dwm[:, :147].reshape(cout, cin, 7, 7) if k == 7 else dwm.view(cout, k, k, cin).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[0]:
            da = torch.empty(m, kk, device=a.device, dtype=torch.float32)
            wt = ConvS2Fn._matrix(weight).t().contiguous()                      # (K, Cout): dA = dy W
            ops.gemm(dy2, ops.SplitWeight(wt), None, da, M=m, N=kk, K=cout, lda=cout, ldw=cout, ldc=kk, relu=False,
                     precision=GRAD_PRECISION if PRECISION.startswith("hybrid16") else PRECISION, short_runs=True)
            dxn = torch.empty(b, h, w, cin, device=a.device, dtype=torch.float32)
            if k == 3:
                check(lib.df_enc_col2im_s2(ptr(da), ptr(dxn), b, h, w, cin, stream()), "df_enc_col2im_s2")
            else:
                dxn.zero_()
                dxn[:, ::2, ::2, :] = da.view(b, (h - 1) // 2 + 1, (w - 1) // 2 + 1, cin)
            dx = dxn.permute(0, 3, 1, 2)
        _s2_sync(f"backward k={k} in={ctx.in_shape}")
        return dx, dw, None


def eligible_s2(m: nn.Conv2d, x: torch.Tensor) -> bool:
    """conv1 7x7/2 (pad 3, 3 input channels), 3x3/2 (pad 1) and 1x1/2 without bias, training mode, on CUDA."""
    if not (STRIDE2_TC and ENABLED and x.is_cuda and x.dtype == torch.float32 and torch.is_grad_enabled()
            and (x.requires_grad or m.weight.requires_grad) and m.stride == (2, 2) and m.groups == 1 and m.bias is None
            and m.dilation == (1, 1) and m.out_channels % 64 == 0):
        return False
    k = m.kernel_size
    if k[0] not in _S2_ONLY:
        return False
    if k == (7, 7):
        return m.padding == (3, 3) and m.in_channels == 3 and not x.requires_grad
    if k == (3, 3):
        return m.padding == (1, 1) and m.in_channels % 64 == 0
    return k == (1, 1) and m.padding == (0, 0) and m.in_channels % 64 == 0


def eligible(m: nn.Conv2d, x: torch.Tensor) -> bool:
    return (ENABLED and x.is_cuda and x.dtype == torch.float32 and torch.is_grad_enabled()
            and (x.requires_grad or m.weight.requires_grad)
            and m.stride == (1, 1) and m.groups == 1 and m.kernel_size in ((1, 1), (3, 3))
            and m.in_channels % 32 == 0 and m.out_channels % 32 == 0
            and m.padding == (m.dilation[0] * (m.kernel_size[0] // 2),) * 2 and m.dilation[0] == m.dilation[1])


def conv2d(m: nn.Conv2d, x: torch.Tensor, act: int = 0, residual: torch.Tensor = None) -> torch.Tensor:
    """act(m(x) + residual), act = 0 | 1 (ReLU).  On the tensor-core training path the activation and the skip connection are
    epilogues of the convolution kernel; anywhere else (inference through the module graph, CPU, ineligible shapes) the plain module
    followed by the torch ops."""
    if eligible(m, x):
        return ConvTCFn.apply(x, m.weight, m.bias, m.dilation[0], act, residual)
    if eligible_s2(m, x) and residual is None:
        return ConvS2Fn.apply(x, m.weight, act)
    y = m(x)
    if residual is not None:
        y = y + residual
    return torch.relu(y) if act else y
