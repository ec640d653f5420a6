"""ctypes loader for the C-ABI library (include/rsn_b200.h).  There is no CPU fallback: every op fails
loudly if the library is missing or the device is not a B200."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RSN_B200_LIB") or os.path.join(_HERE, "librsn_b200.so")   # env override: kernel experiments
_lib = None

P, I64, I32, F32 = c_void_p, c_int64, c_int, c_float
_SIGNATURES = {
    "rsn_version": ([], c_int),
    "rsn_last_error": ([], ctypes.c_char_p),
    "rsn_device_ok": ([], c_int),
    "rsn_sample_spaced": ([P, P, P, P, I64, I32, P, P, I64, I64, P], c_int),
    "rsn_pdf_resample": ([P, I64, P, P, P, P, P, I32, F32, P, P, P, I64, I64, I64, P], c_int),
    "rsn_composite_fwd": ([P, P, P, I64, P, I64, P, P, P, P, I64, I64, P], c_int),
    "rsn_composite_bwd": ([P, P, P, I64, P, I64, P, P, P, P, P, I64, I64, P], c_int),
    "rsn_field_forward": ([P, P, I32, P, P, P, P, I64, I64, P, P, P], c_int),
    "rsn_field_forward_train": ([P, P, I32, P, P, P, P, I64, I64, P, P, P, P, P], c_int),
    "rsn_field_stash_bytes": ([I64], c_int64),
    "rsn_field_normals": ([P, P, P, I64, I64, P, P], c_int),
    "rsn_field_blob_t_bytes": ([], c_int64),
    "rsn_field_backward": ([P, P, I32, P, P, P, P, I64, I64, P, P, P, P, P, P, P], c_int),
    "rsn_field_dy_stash_bytes": ([I64], c_int64),
    "rsn_field_backward_fused": ([P, P, I32, P, P, P, P, I64, I64, P, P, P, P, P, P, P, P, P], c_int),
    "rsn_field_backward_fused_workspace_bytes": ([I64], c_int64),
    "rsn_field_wgrad": ([P, P, I64, P, P], c_int),
    "rsn_field_wgrad_layout": ([P, P, P], c_int),
    "rsn_unpack_grads": ([P, P, P], c_int),
    "rsn_field_flat_layout": ([P], c_int64),
    "rsn_pack_field": ([P, P, P, P, P, P], c_int),
    "rsn_field_blob_bytes": ([], c_int64),
    "rsn_field_bias_count": ([], c_int64),
    "rsn_ipe_freqs": ([P], c_int),
    "rsn_composite16_fwd": ([P, P, P, I64, P, P, P, P, P, P, P, P, I64, I64, P], c_int),
    "rsn_composite16_bwd": ([P, P, P, I64, P, P, P, P, P, P, P, P, P, I64, I64, P], c_int),
    "rsn_reflect_setup": ([P, P, P, P, P, I32, P, P, P, P, P, P, P, I64, P], c_int),
    "rsn_reflect_compose_fwd": ([P, P, P, P, P, I64, P, P, I32, P, I64, I64, P], c_int),
    "rsn_reflect_compose_bwd": ([P, P, P, P, P, I64, P, P, P, P, P, I64, I64, P], c_int),
    "rsn_probe_umma_kmajor": ([P, P, I64, I64, I64, P, P], c_int),
    "rsn_probe_umma_2cta": ([P, P, I64, I64, P, P], c_int),
    "rsn_field_wgrad_finish": ([P, P, P, P, P], c_int),
    "rsn_probe_epilogue": ([I64, I64, I64, P, P], c_int),
    "rsn_probe_tmem_rate": ([I64, I64, I64, I64, P, P], c_int),
    "rsn_probe_umma_ts": ([P, P, I64, I64, P, I64, P, P], c_int),
    "rsn_probe_umma_rate_2cta": ([I64, I64, I64, P, P], c_int),
    "rsn_probe_umma_rate": ([I32, I32, I64, I64, P, P], c_int),
    "rsn_probe_umma_mnmajor": ([P, P, I64, I64, P, P], c_int),
}


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m reflect_sampling_nerf_b200.build` "
                "(there is no CPU or PyTorch fallback for the rsn_b200 kernels)")
        _lib = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.argtypes = argtypes
            fn.restype = restype
    return _lib


def call(name: str, *args) -> None:
    code = getattr(lib(), name)(*args)
    if code != 0:
        raise RuntimeError(f"{name} failed with code {code}: {lib().rsn_last_error().decode()}")


def ptr(t):
    """Device pointer of a (contiguous-as-needed) CUDA tensor, or NULL for None."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("rsn_b200 ops need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream
