"""Colour encoder (ModifiedResnet = BN-free dilated ResNet-18 + PSP pyramid + 3 up-sampling stages, reference:
lib/network.py:27-37, lib/extractors.py:78-124, lib/pspnet.py:7-77) on the tensor cores -- SURVEY.md section 8f row N1.

After the head moved to tcgen05 the torch/cuDNN fp32 encoder was 85-90% of the pose step.  Here every convolution is a
GEMM on the CTA-pair tcgen05 kernel in the same fp32-parity arithmetic as the head (`precision`: "hybrid16s" by default):
  * 3x3 stride-1 convolutions (any dilation): implicit GEMM, the A operand is a 4-D TMA box over the NHWC activation
    shifted by the tap -- zero padding is TMA's out-of-bounds fill, nothing is im2col'ed (df_conv_tc);
  * 1x1 convolutions: plain GEMMs over the (pixels, channels) matrix;
  * the three stride-2 layers (conv1 7x7, layer2.0.conv1 3x3, layer2.0.downsample 1x1) go through small im2col
    buffers (the 1x1 one is the centre tap of the 3x3 one) and the same GEMM;
  * skip connections, biases, ReLU / PReLU are GEMM epilogues;
  * the decoder stages (x2 resize + 3x3 convolution + PReLU) are evaluated at their LOW resolution: a 1x1 convolution commutes
    with the bilinear resize, so x . [W_tap0 .. W_tap8] is one GEMM over a quarter of the pixels and df_enc_upconv_finish sums
    the nine shifted bilinear samples (4x fewer FLOPs on what were the two most expensive layers, no resized maps);
  * the pyramid (PSP) module is folded: 1x1 convolutions commute with the bilinear resize, so the four pooled branches are
    multiplied by (bottleneck slice x stage weight) at 50 cells per crop, resized and summed by one kernel, and enter the
    bottleneck GEMM as its residual operand -- K = 512 instead of 2560 at full resolution and no 2560-wide concat.
Activations are NHWC fp32; the result (B,H,W,32) log-softmax embedding is consumed by df_gather_embedding through
explicit strides.  Inference only: training runs the module graph of lib/pspnet.py / lib/extractors.py, whose convolutions
(lib/conv_tc.py) and element-wise / pooling ops (lib/ew.py) are own kernels with explicit backward passes."""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch

from . import ops
from ._C import check, lib, ptr, stream

K_CONV1 = 160          # 3*7*7 = 147 padded to a multiple of 32
_NARROW_ON_3XTF32 = os.environ.get("DF_HYBRID_NARROW", "0") != "1"     # A/B knob for the rule in PackedEncoder._conv
_CONV1_GATHER = os.environ.get("DF_CONV1_GATHER", "1") == "1"           # A/B knob: conv1 on df_enc_conv1_tc (hybrid16s) vs im2col + GEMM


def _pack_conv(w: torch.Tensor) -> ops.SplitWeight:
    """torch (Cout,Cin,kh,kw) -> (Cout, kh*kw*Cin) tap-major, channels fastest (matches the shifted NHWC boxes)."""
    co, ci, kh, kw = w.shape
    sw = ops.SplitWeight(w.detach().float().permute(0, 2, 3, 1).reshape(co, kh * kw * ci).contiguous())
    sw.split()
    return sw


class PackedEncoder:
    def __init__(self, cnn: torch.nn.Module):
        net = cnn.model.module                      # PSPNet behind the DataParallel-compatible shim
        sd = {k: v.detach() for k, v in net.state_dict().items()}
        dev = next(net.parameters()).device
        self.device = dev
        w1 = torch.zeros(64, K_CONV1, device=dev)
        w1[:, :147] = sd["feats.conv1.weight"].float().reshape(64, 147)
        self.conv1 = ops.SplitWeight(w1)
        self.conv1.split()
        self.blocks = []
        for li, (name, stride, dil) in enumerate((("layer1", 1, 1), ("layer2", 2, 1), ("layer3", 1, 2), ("layer4", 1, 4))):
            for bi in (0, 1):
                pre = f"feats.{name}.{bi}."
                blk = {"c1": _pack_conv(sd[pre + "conv1.weight"]), "c2": _pack_conv(sd[pre + "conv2.weight"]),
                       "dil": 1 if bi == 0 else dil, "stride": stride if bi == 0 else 1, "down": None,
                       "cout": sd[pre + "conv1.weight"].shape[0], "cin": sd[pre + "conv1.weight"].shape[1]}
                if pre + "downsample.0.weight" in sd:
                    blk["down"] = _pack_conv(sd[pre + "downsample.0.weight"])
                self.blocks.append(blk)
        self.bottleneck_b = sd["psp.bottleneck.bias"].float().contiguous()
        # folded pyramid: bottleneck slice x stage convolution, multiplied once in fp64 (see encoder.cu, "folded pyramid")
        wb = sd["psp.bottleneck.weight"].double().reshape(1024, 2560)
        self.pyramid = []
        for i in range(4):
            sw = ops.SplitWeight((wb[:, 512 * i:512 * (i + 1)] @ sd[f"psp.stages.{i}.1.weight"].double().reshape(512, 512)).float().contiguous())
            sw.split()
            self.pyramid.append(sw)
        self.bottleneck_f = ops.SplitWeight(wb[:, 2048:].float().contiguous())
        self.bottleneck_f.split()
        self.ups = []
        for u in ("up_1", "up_2", "up_3"):
            w = sd[f"{u}.conv.1.weight"]
            co, ci = w.shape[0], w.shape[1]
            # "wz": the nine taps stacked along the OUTPUT axis, (9*Cout, Cin) -- the low-resolution GEMM of the folded stage
            wz = ops.SplitWeight(w.detach().float().permute(2, 3, 0, 1).reshape(9 * co, ci).contiguous())
            wz.split()
            self.ups.append({"wz": wz, "b": sd[f"{u}.conv.1.bias"].float().contiguous(),
                             "a": sd[f"{u}.conv.2.weight"].float().contiguous(), "cout": co})
        self.ups[2]["w"] = _pack_conv(sd["up_3.conv.1.weight"])       # (Cout, 9*Cin): the sparse tail's patch GEMM
        self.final_w = sd["final.0.weight"].float().reshape(32, 64).contiguous()
        self.final_b = sd["final.0.bias"].float().contiguous()
        self._ws: Dict[Tuple[int, int, int], dict] = {}

    # ---- scratch: one set of NHWC buffers per bucket shape ----
    def _workspace(self, b, H, W):
        key = (b, H, W)
        ws = self._ws.get(key)
        if ws is None:
            f = dict(device=self.device, dtype=torch.float32)
            H2, W2 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
            H4, W4 = (H2 - 1) // 2 + 1, (W2 - 1) // 2 + 1
            H8, W8 = (H4 - 1) // 2 + 1, (W4 - 1) // 2 + 1
            ws = {"dims": (H2, W2, H4, W4, H8, W8)}
            ws["a0"] = None                                  # conv1 patch matrix: only the modes without the gathering kernel need it
            ws["c1"] = torch.empty(b, H2, W2, 64, **f)
            ws["l1"] = [torch.empty(b, H4, W4, 64, **f) for _ in range(3)]
            ws["a2"] = torch.empty(b * H8 * W8, 9 * 64, **f)
            ws["l8"] = {c: [torch.empty(b, H8, W8, c, **f) for _ in range(3)] for c in (128, 256, 512)}
            ws["bott"] = torch.empty(b, H8, W8, 1024, **f)
            ws["pyr_pool"] = torch.empty(50 * b, 512, **f)
            ws["pyr_y"] = torch.empty(50 * b, 1024, **f)
            ws["pyr_sum"] = torch.empty(b, H8, W8, 1024, **f)
            hs = [(2 * H8, 2 * W8), (4 * H8, 4 * W8), (8 * H8, 8 * W8)]
            ws["hs"] = hs
            # decoder stages are allocated on first use: the sparse tail (forward_points) never needs the
            # full-resolution third stage (1 GB per bucket of 64 crops at 160x160)
            ws["up_in"], ws["up_out"], ws["feat"] = [None] * 3, [None] * 3, None
            self._ws[key] = ws
        return ws

    def _stage_buffers(self, ws, i, b):
        if ws["up_in"][i] is None:
            f = dict(device=self.device, dtype=torch.float32)
            (h, w), cout = ws["hs"][i], (256, 64, 64)[i]
            ws["up_in"][i] = torch.empty(b, h // 2, w // 2, 9 * cout, **f)      # Z: the nine tap products at the low resolution
            ws["up_out"][i] = torch.empty(b, h, w, cout, **f)
        return ws["up_in"][i], ws["up_out"][i]

    # ---- kernels ----
    @staticmethod
    def _conv(x, w: ops.SplitWeight, out, *, taps, dil=1, bias=None, residual=None, prelu=None, act=1, mode=1, cin=None,
              ldx=None, ldy=None, cout=None):
        b, h, wd = x.shape[0], x.shape[1], x.shape[2]
        cin = x.shape[3] if cin is None else cin
        ldx = x.stride(2) if ldx is None else ldx
        cout = out.shape[3] if cout is None else cout
        ldy = out.stride(2) if ldy is None else ldy
        if mode in (3, 4) and cout <= 64 and _NARROW_ON_3XTF32:
            mode = 1        # 64-wide tiles are paced by the A stagers, where the hybrid split costs more (measured +9%); not so in
                            # hybrid16s (cheaper split, four TMEM A stages: the 64-channel layers are 15-20% faster than in 3xtf32)
        hi, lo = w.operands(mode)
        check(lib.df_conv_tc(ptr(x), b, h, wd, cin, ldx, ptr(hi), ptr(lo), taps, dil, ptr(bias), ptr(residual),
                             0 if residual is None else residual.stride(2), ptr(prelu), act, ptr(out), ldy, cout, mode,
                             stream()), "df_conv_tc")

    def forward(self, img: torch.Tensor, precision: str = "3xtf32") -> torch.Tensor:
        """img (b,3,H,W) fp32 CUDA (NCHW, as the reference feeds it) -> (b,H,W,32) NHWC log-softmax embedding.
        H and W must be multiples of 8 (the reference's crops are multiples of 40)."""
        x, ws, mode = self._trunk(img, precision, stages=3)
        b, h, w = x.shape[0], x.shape[1], x.shape[2]
        rows = b * h * w
        if ws["feat"] is None:
            ws["feat"] = torch.empty(b, h, w, 32, device=self.device, dtype=torch.float32)
        ops.gemm(x, self.final_w, self.final_b, ws["feat"], M=rows, N=32, K=64, lda=64, ldw=64, ldc=32, relu=False)
        check(lib.df_enc_log_softmax32(ptr(ws["feat"]), rows, stream()), "df_enc_log_softmax32")
        return ws["feat"]

    def forward_points(self, img: torch.Tensor, choose: torch.Tensor, emb_pm_out: torch.Tensor,
                       precision: str = "3xtf32") -> torch.Tensor:
        """The embedding only where the head reads it: emb_pm_out (b*N,32) = embedding at pixels choose (b,N).

        The head gathers N = 500 of the H*W (6 400 .. 25 600) pixels, so the last decoder stage -- x2 up-sampling, the
        3x3 64->64 convolution + PReLU, the 1x1 64->32 convolution and the channel log-softmax (lib/pspnet.py:34-37,
        :53-56, lib/network.py:98-102) -- is evaluated on gathered 3x3 patches of those pixels only: identical values,
        2-8% of the work and none of the full-resolution intermediates."""
        x, ws, mode = self._trunk(img, precision, stages=2)             # (b, H/2, W/2, 64)
        b, h, w = x.shape[0], x.shape[1], x.shape[2]
        choose = ops.i64c(choose).view(b, -1)
        n = choose.shape[1]
        rows = b * n
        key = ("pts", n)
        if key not in ws:
            f = dict(device=self.device, dtype=torch.float32)
            ws[key] = (torch.empty(rows, 9 * 64, **f), torch.empty(rows, 64, **f))
        patches, act = ws[key]
        s = stream()
        check(lib.df_enc_gather_up_patches(ptr(x), ptr(choose), ptr(patches), b, n, h, w, 64, s), "df_enc_gather_up_patches")
        up = self.ups[2]
        # the gathered patches as a 1 x rows "image" with 576 channels: a 1-tap convolution = GEMM with the PReLU epilogue
        self._conv(patches.view(1, 1, rows, 576), up["w"], act.view(1, 1, rows, 64), taps=1, bias=up["b"], prelu=up["a"], act=2, mode=mode)
        ops.gemm(act, self.final_w, self.final_b, emb_pm_out, M=rows, N=32, K=64, lda=64, ldw=64, ldc=32, relu=False)
        check(lib.df_enc_log_softmax32(ptr(emb_pm_out), rows, s), "df_enc_log_softmax32")
        return emb_pm_out

    def _trunk(self, img: torch.Tensor, precision: str, stages: int):
        """Everything up to and including `stages` of the three up-sampling stages.  Returns (activation, workspace, mode)."""
        if precision not in ("3xtf32", "tf32", "hybrid", "hybrid16", "hybrid16s"):
            raise ValueError("the tensor-core encoder runs in '3xtf32' / 'hybrid' / 'hybrid16' / 'hybrid16s' (fp32 parity) or 'tf32'")
        mode = ops.PRECISIONS[precision]
        img = ops.f32c(img)
        b, _, H, W = img.shape
        if H % 8 or W % 8:
            raise ValueError("encoder: crop height / width must be multiples of 8")
        ws = self._workspace(b, H, W)
        H2, W2, H4, W4, H8, W8 = ws["dims"]
        s = stream()
        # conv1 7x7/2 + ReLU, 3x3/2 max pool
        rows = b * H2 * W2
        if mode == 6 and _CONV1_GATHER:
            # hybrid16s: the kernel's stagers gather the 7x7 patches from the image themselves -- no patch matrix (0.58 GB per step)
            planes, scale = self.conv1.planes16s()
            check(lib.df_enc_conv1_tc(ptr(img), b, H, W, ptr(planes), ptr(scale), ptr(ws["c1"]), 64, 64, 1, s), "df_enc_conv1_tc")
        else:
            if ws["a0"] is None:
                ws["a0"] = torch.empty(rows, K_CONV1, device=self.device, dtype=torch.float32)
            check(lib.df_enc_im2col_conv1(ptr(img), ptr(ws["a0"]), b, H, W, K_CONV1, s), "df_enc_im2col_conv1")
            ops.gemm(ws["a0"], self.conv1, None, ws["c1"], M=rows, N=64, K=K_CONV1, lda=K_CONV1, ldw=K_CONV1, ldc=64, relu=True,
                     precision=precision)
        x = ws["l1"][0]
        check(lib.df_enc_maxpool(ptr(ws["c1"]), ptr(x), b, H2, W2, 64, s), "df_enc_maxpool")
        # residual stages
        free = [ws["l1"][1], ws["l1"][2]]
        for bi, blk in enumerate(self.blocks):
            cout = blk["cout"]
            if blk["stride"] == 2:
                # layer2.0: 3x3/2 and the 1x1/2 projection both read the im2col'ed patches (projection = centre tap)
                check(lib.df_enc_im2col_s2(ptr(x), ptr(ws["a2"]), b, H4, W4, 64, s), "df_enc_im2col_s2")
                r8 = b * H8 * W8
                bufs = ws["l8"][128]
                t, skip, out = bufs[0], bufs[1], bufs[2]
                ops.gemm(ws["a2"], blk["c1"], None, t, M=r8, N=128, K=576, lda=576, ldw=576, ldc=128, relu=True,
                         precision=precision)
                ops.gemm(ws["a2"][:, 256:], blk["down"], None, skip, M=r8, N=128, K=64, lda=576, ldw=64, ldc=128,
                         relu=False, precision=precision)
                self._conv(t, blk["c2"], out, taps=9, dil=1, residual=skip, act=1, mode=mode)
                x, free = out, [t, skip]
                continue
            if cout != x.shape[3]:                                   # first block of layer3 / layer4: wider, projected skip
                bufs = ws["l8"][cout]
                t, skip, out = bufs[0], bufs[1], bufs[2]
                self._conv(x, blk["down"], skip, taps=1, act=0, mode=mode)
                self._conv(x, blk["c1"], t, taps=9, dil=blk["dil"], act=1, mode=mode)
                self._conv(t, blk["c2"], out, taps=9, dil=blk["dil"], residual=skip, act=1, mode=mode)
                x, free = out, [t, skip]
                continue
            t, out = free[0], free[1]
            self._conv(x, blk["c1"], t, taps=9, dil=blk["dil"], act=1, mode=mode)
            self._conv(t, blk["c2"], out, taps=9, dil=blk["dil"], residual=x, act=1, mode=mode)
            x, free = out, [t, x]
        # pyramid pooling (lib/pspnet.py:17-24): relu(Wb . cat[up(S_s . pool_s(f)) for s in 1,2,3,6; f] + bias)
        #   = relu(Wb_f . f + bias + sum_s up((Wb_s S_s) . pool_s(f))): the pooled branches are evaluated at 50 cells per crop
        #   and only the 512 channels of f go through the full-resolution GEMM (K = 512 instead of 2560)
        feats = x
        check(lib.df_enc_pyramid_pool(ptr(feats), 512, ptr(ws["pyr_pool"]), b, H8, W8, 512, s), "df_enc_pyramid_pool")
        row = 0
        for i, sz in enumerate((1, 2, 3, 6)):
            r = b * sz * sz
            ops.gemm(ws["pyr_pool"][row:], self.pyramid[i], None, ws["pyr_y"][row:], M=r, N=1024, K=512, lda=512, ldw=512,
                     ldc=1024, relu=False, precision=precision)
            row += r
        check(lib.df_enc_pyramid_sum(ptr(ws["pyr_y"]), ptr(ws["pyr_sum"]), 1024, b, H8, W8, 1024, s), "df_enc_pyramid_sum")
        self._conv(feats, self.bottleneck_f, ws["bott"], taps=1, bias=self.bottleneck_b, residual=ws["pyr_sum"], act=1, mode=mode)
        # three decoder stages (x2 bilinear resize, align_corners; 3x3 convolution; PReLU), each evaluated at its LOW resolution:
        # conv3x3(up(x)) = sum_tap shift_tap(up(W_tap . x)) -- one GEMM x . [W_0..W_8] over a quarter of the pixels, then
        # df_enc_upconv_finish sums the nine shifted bilinear samples, adds the bias and applies PReLU (4x fewer FLOPs, no resized map)
        x, (h, w) = ws["bott"], (H8, W8)
        for i, up in enumerate(self.ups[:stages]):
            c, cout = x.shape[3], up["cout"]
            z, up_out = self._stage_buffers(ws, i, b)
            ops.gemm(x, up["wz"], None, z, M=b * h * w, N=9 * cout, K=c, lda=c, ldw=c, ldc=9 * cout, relu=False, precision=precision)
            check(lib.df_enc_upconv_finish(ptr(z), 9 * cout, ptr(up["b"]), ptr(up["a"]), ptr(up_out), cout, b, h, w, cout, s),
                  "df_enc_upconv_finish")
            x, (h, w) = up_out, ws["hs"][i]
        return x, ws, mode
