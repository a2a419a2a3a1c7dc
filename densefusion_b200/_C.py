"""ctypes binding of libdensefusion_b200.so (the C ABI declared in include/densefusion_b200.h).

There is NO fallback: if the library is missing or a symbol is absent the import fails loudly, and
every wrapper raises on a non-zero status.  torch is used for device memory and streams only."""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdensefusion_b200.so")

_p = ctypes.c_void_p
_i = ctypes.c_int
_ll = ctypes.c_longlong
_ull = ctypes.c_ulonglong
_f = ctypes.c_float

# name -> argtypes ; must list EVERY symbol of include/densefusion_b200.h (tests/test_abi.py checks)
SIGNATURES = {
    "df_abi_version": [],
    "df_features": [],
    "df_probe_ffma": [_p, _i, _i, _p],
    "df_knn": [_p, _p, _p, _i, _i, _i, _i, _i, _p],
    "df_loss_forward": [_p, _p, _p, _p, _p, _p, _p, _p, _ull, _i, _f, _i, _i, _i, _i,
                        _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "df_loss_backward": [_p, _p, _p, _p, _p, _p, _p, _p, _f, _i, _i, _p, _p, _p, _p],
    "df_gemm_fp32": [_p, _i, _p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _p, _p],
    "df_gemm_rows_per_pool_tile": [],
    "df_gemm_tc": [_p, _i, _p, _p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _ll, _ll, _ll, _p, _i, _i, _p],
    "df_split_tf32": [_p, _p, _p, _ll, _p],
    "df_pack_bf16_pairs": [_p, _p, _ll, _i, _p],
    "df_pack_f16_pairs": [_p, _p, _p, _ll, _i, _p],
    "df_pack_f16s": [_p, _p, _p, _ll, _i, _p],
    "df_tc_trace_read": [_p, _i],
    "df_enc_conv1_tc": [_p, _i, _i, _i, _p, _p, _p, _i, _i, _i, _p],
    "df_ew_relu_mask": [_p, _p, _p, _i, _i, _ll, _p],
    "df_ew_maxpool_backward": [_p, _p, _p, _i, _i, _i, _i, _p],
    "df_ew_pyramid_pool_backward": [_p, _p, _i, _i, _i, _i, _i, _p],
    "df_ew_log_softmax32_backward": [_p, _p, _p, _ll, _p],
    "df_ew_prelu": [_p, _p, _p, _ll, _p],
    "df_ew_prelu_scratch_floats": [],
    "df_ew_prelu_backward": [_p, _p, _p, _p, _p, _p, _ll, _p],
    "df_ew_dropout_mask": [_p, _i, _f, _p, _p],
    "df_ew_scale_bc": [_p, _p, _p, _i, _ll, _i, _p],
    "df_ew_copy2d": [_p, _i, _p, _i, _ll, _i, _p],
    "df_ew_add": [_p, _p, _p, _ll, _p],
    "df_pack_conv_weight": [_p, _p, _p, _p, _i, _i, _i, _i, _p],
    "df_pack_conv_weight16": [_p, _p, _p, _i, _i, _i, _i, _p],
    "df_conv_wgrad_tc": [_p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p],
    "df_conv_wgrad_scratch_floats": [_i, _i, _i, _i, _i, _i, _i],
    "df_gemm_dgrad_fp32": [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _ll, _ll, _ll, _p, _i, _p],
    "df_gemm_wgrad_fp32": [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _ll, _ll, _p],
    "df_reduce_partials": [_p, _i, _ll, _p, _i, _p],
    "df_colsum_rows": [_p, _i, _i, _i, _i, _p, _i, _p],
    "df_relu_mask_inplace": [_p, _p, _i, _i, _ll, _p],
    "df_pool_backward": [_p, _p, _p, _i, _i, _ll, _p],
    "df_select_out_backward": [_p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _i, _ll, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "df_gather_embedding_backward": [_p, _p, _p, _ll, _ll, _ll, _i, _i, _i, _p],
    "df_adam_step": [_p, _p, _p, _p, _ll, _f, _f, _f, _f, _i, _p],
    "df_adam_step_dev": [_p, _p, _p, _p, _ll, _f, _f, _f, _f, _p, _p],
    "df_conv_tc_macs": [_i, _i, _i, _i, _i, _i, _i],
    "df_conv_tc_schedule": [_i, _i, _i, _i, _i, _i, _i, _p],
    "df_gemm_tc_plan": [_i, _i, _i, _i, _i, _i, _i, _p],
    "df_conv_tc": [_p, _i, _i, _i, _i, _i, _p, _p, _i, _i, _p, _p, _i, _p, _i, _p, _i, _i, _i, _p],
    "df_enc_im2col_conv1": [_p, _p, _i, _i, _i, _i, _p],
    "df_enc_maxpool": [_p, _p, _i, _i, _i, _i, _p],
    "df_enc_im2col_s2": [_p, _p, _i, _i, _i, _i, _p],
    "df_enc_col2im_s2": [_p, _p, _i, _i, _i, _i, _p],
    "df_enc_adaptive_avgpool": [_p, _i, _p, _i, _i, _i, _i, _i, _p],
    "df_enc_pyramid_pool": [_p, _i, _p, _i, _i, _i, _i, _p],
    "df_enc_pyramid_sum": [_p, _p, _i, _i, _i, _i, _i, _p],
    "df_enc_upconv_finish": [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "df_enc_upsample": [_p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "df_enc_upsample_backward": [_p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "df_enc_log_softmax32": [_p, _ll, _p],
    "df_enc_gather_up_patches": [_p, _p, _p, _i, _i, _i, _i, _i, _p],
    "df_build_crops": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, ctypes.c_uint, _p, _p, _p, _p, _p],
    "df_upsample_bilinear": [_p, _p, _ll, _i, _i, _i, _i, _i, _p],
    "df_gather_embedding": [_p, _p, _p, _p, _ll, _ll, _ll, _i, _i, _i, _p],
    "df_xyz_conv": [_p, _p, _p, _p, _i, _ll, _p],
    "df_pool_finish": [_p, _p, _i, _i, _i, _i, _p],
    "df_select_out": [_p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _ll, _p, _p, _p, _p],
    "df_select_pose": [_p, _p, _p, _p, _i, _i, _p, _p, _p],
    "df_cloud_transform": [_p, _p, _p, _i, _i, _p],
    "df_pose_compose": [_p, _p, _p, _i, _p],
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is mandatory (no CPU / eager fallback). "
            "Build it with `python -m densefusion_b200.build` or __graft_entry__.build().")
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing -> loud
        fn.argtypes = args
        fn.restype = ctypes.c_longlong if name in ("df_conv_wgrad_scratch_floats", "df_probe_ffma", "df_conv_tc_macs") else ctypes.c_int
    return lib


class _CountingLib:
    """Forwards to the CDLL and counts kernel-launching entry points (bench.py reports `gpu_launches`)."""
    _NO_LAUNCH = ("df_abi_version", "df_features", "df_gemm_rows_per_pool_tile", "df_conv_wgrad_scratch_floats", "df_conv_tc_macs", "df_conv_tc_schedule", "df_gemm_tc_plan",
                  "df_ew_prelu_scratch_floats", "df_tc_trace_read")

    _TIMED = ("df_gemm_tc", "df_conv_tc")
    # entry points that launch more than one kernel: the pyramid pool runs its 1x1 / 2x2 stages and its 3x3 / 6x6 stages as two kernels
    # (csrc/encoder.cu::df_enc_pyramid_pool; DF_ENC_V1 / DF_ENC_POOL_SPLIT select single-kernel forms)
    _KERNELS = {"df_enc_pyramid_pool": 2 if (os.environ.get("DF_ENC_V1", "0") in ("", "0")
                                             and os.environ.get("DF_ENC_POOL_SPLIT", "1") == "1") else 1}

    def __init__(self, cdll):
        self._cdll = cdll
        self.launches = 0
        self.timer = None          # bench.py: a list -> every tensor-core launch is bracketed by CUDA events and appended
        for name in SIGNATURES:
            fn = getattr(cdll, name)
            if name in self._NO_LAUNCH:
                setattr(self, name, fn)
            else:
                setattr(self, name, self._counted(name, fn))

    def _counted(self, name, fn):
        timed = name in self._TIMED

        kernels = self._KERNELS.get(name, 1)

        def call(*args):
            self.launches += kernels
            t = self.timer
            if t is None or not timed:
                return fn(*args)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            t.append((name, args, e0, e1, rc))
            return rc
        return call


lib = _CountingLib(_load())


class DFError(RuntimeError):
    pass


def check(status: int, what: str) -> None:
    if status == 0:
        return
    if status == -1:
        raise DFError(f"{what}: invalid argument")
    if status == -2:
        raise DFError(f"{what}: unsupported configuration")
    raise DFError(f"{what}: CUDA error {status}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Kernels launch on the CURRENT device's current stream (see stream()), so a
    CUDA tensor living on another device is an error here rather than an illegal address inside the kernel."""
    if t is None:
        return None
    if t.is_cuda and t.device.index != torch.cuda.current_device():
        raise DFError(f"tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                      "call torch.cuda.set_device(...) / use torch.cuda.device(...) around densefusion_b200 calls")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise DFError("densefusion_b200 kernels need CUDA tensors (there is no CPU path)")


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32 + contiguous (a no-op view when already so)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def i64c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.int64:
        t = t.long()
    return t if t.is_contiguous() else t.contiguous()
