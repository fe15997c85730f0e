"""ctypes binding of libnesvor_b200.so (the C ABI declared in include/nesvor_b200.h).

There is deliberately NO fallback: if the shared object is missing or a call fails, a RuntimeError
is raised.  PyTorch is used only as the owner of device memory and streams.
"""
import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libnesvor_b200.so")
NSV_MAX_LEVELS = 32

_lib: Optional[ctypes.CDLL] = None


class GridMeta(ctypes.Structure):
    """struct nsv_grid_meta"""

    _fields_ = [
        ("n_levels", ctypes.c_int32),
        ("n_features", ctypes.c_int32),
        ("scale", ctypes.c_float * NSV_MAX_LEVELS),
        ("res", ctypes.c_uint32 * NSV_MAX_LEVELS),
        ("size", ctypes.c_uint32 * NSV_MAX_LEVELS),
        ("offset", ctypes.c_uint32 * (NSV_MAX_LEVELS + 1)),
        ("hashed", ctypes.c_uint32 * NSV_MAX_LEVELS),
    ]


class InrConfig(ctypes.Structure):
    """struct nsv_inr_config"""

    _fields_ = [
        ("grid", GridMeta),
        ("width", ctypes.c_int32),
        ("depth", ctypes.c_int32),
        ("n_features_z", ctypes.c_int32),
        ("n_features_slice", ctypes.c_int32),
        ("n_levels_bias", ctypes.c_int32),
        ("pixel_variance", ctypes.c_int32),
        ("slice_variance", ctypes.c_int32),
        ("slice_scale", ctypes.c_int32),
        ("pose_grad", ctypes.c_int32),
        ("image_reg", ctypes.c_int32),
        ("delta", ctypes.c_float),
        ("w_image", ctypes.c_float),
        ("w_bias", ctypes.c_float),
        ("bbox_lo", ctypes.c_float * 3),
        ("bbox_hi", ctypes.c_float * 3),
        ("grad_scale", ctypes.c_float),
    ]


class InrParams(ctypes.Structure):
    """struct nsv_inr_params"""

    _fields_ = [
        ("table_f16", ctypes.c_void_p),
        ("mlp_f16", ctypes.c_void_p),
        ("axisangle", ctypes.c_void_p),
        ("psf_sigma", ctypes.c_void_p),
        ("slice_embedding", ctypes.c_void_p),
        ("logit_coef", ctypes.c_void_p),
        ("log_var_slice", ctypes.c_void_p),
        ("n_slices", ctypes.c_int32),
    ]


class InrGrads(ctypes.Structure):
    """struct nsv_inr_grads"""

    _fields_ = [
        ("table", ctypes.c_void_p),
        ("mlp", ctypes.c_void_p),
        ("axisangle", ctypes.c_void_p),
        ("slice_embedding", ctypes.c_void_p),
        ("slice_scale_c", ctypes.c_void_p),
        ("log_var_slice", ctypes.c_void_p),
        ("losses", ctypes.c_void_p),
    ]


def lib() -> ctypes.CDLL:
    """Loads the native library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"nesvor_b200: native library {LIB_PATH} not found. Build it with "
                "`python -m nesvor_b200.csrc.build` (or __graft_entry__.build()); there is no CPU/PyTorch fallback."
            )
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.nsv_last_error_string.restype = ctypes.c_char_p
        _lib.nsv_build_arch.restype = ctypes.c_char_p
        _lib.nsv_grid_meta_init.restype = ctypes.c_int64
        if hasattr(_lib, "nsv_inr_mlp_layout"):
            _lib.nsv_inr_mlp_layout.restype = ctypes.c_int64
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().nsv_last_error_string().decode()
        raise RuntimeError(f"nesvor_b200 native call failed{(' in ' + what) if what else ''}: rc={rc} {msg}")


def require_cuda(name: str, t: Optional[torch.Tensor], dtype=None) -> None:
    """Mirror of the reference's CHECK_INPUT (slice_acq_cuda.cpp:57-59): CUDA + contiguous."""
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must have dtype {dtype}, got {t.dtype}")


def ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    if t is None or t.numel() == 0:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream(device=None) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def make_grid_meta(n_levels: int, n_features: int, log2_hashmap_size: int, base_resolution: int, per_level_scale: float):
    """Returns (GridMeta, total number of table entries)."""
    m = GridMeta()
    total = lib().nsv_grid_meta_init(
        ctypes.byref(m), ctypes.c_int(n_levels), ctypes.c_int(n_features), ctypes.c_int(log2_hashmap_size),
        ctypes.c_int(base_resolution), ctypes.c_float(per_level_scale))
    if total < 0:
        check(int(total), "nsv_grid_meta_init")
    return m, int(total)


FUSED_IMPLS = {"auto": 0, "mma": 1, "tcgen05": 2, "ws": 3}


def set_fused_impl(name: str) -> None:
    """Selects the implementation of kernel A: "auto" (tcgen05/TMEM when instantiated, else mma.sync),
    "mma" (mma.sync fragments) or "tcgen05" (fail with FusedUnsupported if not instantiated)."""
    check(lib().nsv_set_fused_impl(ctypes.c_int(FUSED_IMPLS[name])), "nsv_set_fused_impl")


def set_fused_tuning(agg_max_entries: int = -1, fast_path: int = -1) -> None:
    """Tuning / test hook of kernel A's gather-scatter loops (see include/nesvor_b200.h)."""
    check(lib().nsv_set_fused_tuning(ctypes.c_int64(agg_max_entries), ctypes.c_int(fast_path)), "nsv_set_fused_tuning")
