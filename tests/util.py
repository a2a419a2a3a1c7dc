"""Shared helpers for the parity tests."""
import ctypes
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    """max |a-b| / max |b|  (the 1e-4 'relative' bound of BASELINE.json is checked in this norm)."""
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


def rel_elementwise(a, b, floor=1e-3):
    """max_i |a_i - b_i| / max(|b_i|, floor * max|b|): element-wise relative error with an absolute floor, so that outputs near
    zero (e.g. pred_t offsets) are not judged against the largest entry of the tensor only."""
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), floor * (np.max(np.abs(b)) + 1e-30))
    return float(np.max(np.abs(a - b) / den))


def build_nets(num_points, num_obj, seed, device="cuda"):
    """Drop-in modules loaded with the by-name synthetic weights (same values the golden script loaded into
    the reference modules)."""
    from densefusion_b200 import synth
    from densefusion_b200.lib.network import PoseNet, PoseRefineNet
    est, ref = PoseNet(num_points, num_obj), PoseRefineNet(num_points, num_obj)
    est_sd = synth.synth_state_dict(synth.shapes_of(est), seed)
    ref_sd = synth.synth_state_dict(synth.shapes_of(ref), seed + 1)
    est.load_state_dict(est_sd)
    ref.load_state_dict(ref_sd)
    est.eval().requires_grad_(False)
    ref.eval().requires_grad_(False)
    if device is not None:
        est.to(device)
        ref.to(device)
    return est, ref, est_sd, ref_sd


_REFKNN = None


def reference_knn_gpu(ref: torch.Tensor, query: torch.Tensor, k: int = 1) -> torch.Tensor:
    """The reference's OWN CUDA kernels (oracle/_ref/libknn_reference.so, compiled from
    /root/reference/lib/knn/src/knn_cuda_kernel.cu) on (D,R) / (D,Q) CUDA tensors -> (k,Q) int64, 1-based.
    Queries are chunked so that R*Qc < 2^31 and the scratch stays below 2 GB."""
    global _REFKNN
    if _REFKNN is None:
        path = os.path.join(ROOT, "oracle", "_ref", "libknn_reference.so")
        if not os.path.exists(path):
            return None
        lib = ctypes.CDLL(path)
        lib.knn_device.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.knn_device.restype = None
        _REFKNN = lib
    D, R = ref.shape
    Q = query.shape[1]
    out = torch.empty(k, Q, dtype=torch.int64, device=ref.device)
    qc_max = max(256, min(Q, (1 << 29) // R))
    scratch = torch.empty(R * qc_max, dtype=torch.float32, device=ref.device)
    stream = torch.cuda.current_stream().cuda_stream
    for q0 in range(0, Q, qc_max):
        qc = min(qc_max, Q - q0)
        qchunk = query[:, q0:q0 + qc].contiguous()
        ochunk = torch.empty(k, qc, dtype=torch.int64, device=ref.device)
        _REFKNN.knn_device(ref.data_ptr(), R, qchunk.data_ptr(), qc, D, k, scratch.data_ptr(), ochunk.data_ptr(), stream)
        out[:, q0:q0 + qc] = ochunk
    torch.cuda.synchronize()
    return out
