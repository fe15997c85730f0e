"""slice_acquisition / slice_acquisition_adjoint on the native library (kernel B).

Mirrors nesvor/slice_acquisition/slice_acq.py:22-211: same autograd Functions, same positional
9-argument wrappers, same saved tensors and the same `need_weight` / `equalize` behaviour, on top
of the C ABI (include/nesvor_b200.h) instead of the pybind module nesvor.slice_acq_cuda
(slice_acq_cuda.cpp:156-161).  The four functions `forward`, `backward`, `adjoint_forward`,
`adjoint_backward` below have the pybind module's signatures and return lists of tensors.
Inputs must be contiguous CUDA tensors (RuntimeError otherwise, like CHECK_INPUT); absent masks are
None or empty tensors.  float32 and float64 are supported; launches use torch's current stream.
"""
import ctypes

import torch
from torch.autograd import Function

from .. import _lib


def _suffix(t: torch.Tensor) -> str:
    if t.dtype == torch.float32:
        return "f32"
    if t.dtype == torch.float64:
        return "f64"
    raise RuntimeError(f"slice_acquisition supports float32/float64 volumes, got {t.dtype}")


def _real(t: torch.Tensor, v: float):
    return ctypes.c_float(v) if t.dtype == torch.float32 else ctypes.c_double(v)


def _mask(name, m):
    if m is None or m.numel() == 0:
        return None
    _lib.require_cuda(name, m, torch.bool)
    return m


def _dims(vol_shape, psf, n, slice_shape):
    vals = [*vol_shape, *psf.shape, n, *slice_shape]
    return [ctypes.c_int(int(v)) for v in vals]


def forward(transforms, vol, vol_mask, slices_mask, psf, slice_shape, res_slice, need_weight, interp_psf):
    for name, t in (("transforms", transforms), ("vol", vol), ("psf", psf)):
        _lib.require_cuda(name, t, vol.dtype)
    vol_mask, slices_mask = _mask("vol_mask", vol_mask), _mask("slices_mask", slices_mask)
    n = transforms.shape[0]
    slices = torch.zeros((n, 1, int(slice_shape[0]), int(slice_shape[1])), dtype=vol.dtype, device=vol.device)
    weight = torch.zeros_like(slices) if need_weight else None
    with torch.cuda.device(vol.device):
        rc = getattr(_lib.lib(), "nsv_slice_acq_forward_" + _suffix(vol))(
            _lib.ptr(transforms), _lib.ptr(vol), _lib.ptr(vol_mask), _lib.ptr(slices_mask), _lib.ptr(psf),
            _lib.ptr(slices), _lib.ptr(weight), *_dims(vol.shape[-3:], psf, n, slice_shape),
            _real(vol, res_slice), ctypes.c_int(int(interp_psf)), _lib.stream(vol.device))
    _lib.check(rc, "nsv_slice_acq_forward")
    return [slices, weight] if need_weight else [slices]


def backward(transforms, vol, vol_mask, psf, grad_slices, slices_mask, res_slice, interp_psf, need_vol_grad,
             need_transforms_grad):
    for name, t in (("transforms", transforms), ("vol", vol), ("psf", psf), ("grad_slices", grad_slices)):
        _lib.require_cuda(name, t, vol.dtype)
    vol_mask, slices_mask = _mask("vol_mask", vol_mask), _mask("slices_mask", slices_mask)
    n = transforms.shape[0]
    grad_vol = torch.zeros_like(vol) if need_vol_grad else None
    grad_tf = torch.zeros_like(transforms) if need_transforms_grad else None
    with torch.cuda.device(vol.device):
        rc = getattr(_lib.lib(), "nsv_slice_acq_backward_" + _suffix(vol))(
            _lib.ptr(transforms), _lib.ptr(vol), _lib.ptr(vol_mask), _lib.ptr(psf), _lib.ptr(grad_slices),
            _lib.ptr(slices_mask), _lib.ptr(grad_vol), _lib.ptr(grad_tf),
            *_dims(vol.shape[-3:], psf, n, grad_slices.shape[-2:]), _real(vol, res_slice),
            ctypes.c_int(int(interp_psf)), _lib.stream(vol.device))
    _lib.check(rc, "nsv_slice_acq_backward")
    return [grad_vol, grad_tf]


def adjoint_forward(transforms, psf, slices, slices_mask, vol_mask, vol_shape, res_slice, interp_psf, equalize):
    for name, t in (("transforms", transforms), ("psf", psf), ("slices", slices)):
        _lib.require_cuda(name, t, slices.dtype)
    vol_mask, slices_mask = _mask("vol_mask", vol_mask), _mask("slices_mask", slices_mask)
    n = transforms.shape[0]
    vol = torch.zeros((1, 1) + tuple(int(v) for v in vol_shape), dtype=slices.dtype, device=slices.device)
    vol_weight = torch.zeros_like(vol) if equalize else None
    with torch.cuda.device(slices.device):
        rc = getattr(_lib.lib(), "nsv_slice_acq_adjoint_forward_" + _suffix(slices))(
            _lib.ptr(transforms), _lib.ptr(psf), _lib.ptr(slices), _lib.ptr(slices_mask), _lib.ptr(vol_mask),
            _lib.ptr(vol), _lib.ptr(vol_weight), *_dims(vol_shape, psf, n, slices.shape[-2:]),
            _real(slices, res_slice), ctypes.c_int(int(interp_psf)), ctypes.c_int(int(equalize)),
            _lib.stream(slices.device))
    _lib.check(rc, "nsv_slice_acq_adjoint_forward")
    return [vol, vol_weight]


def adjoint_backward(transforms, grad_vol, vol_weight, vol_mask, psf, slices, slices_mask, vol, res_slice, interp_psf,
                     equalize, need_slices_grad, need_transforms_grad):
    """NB: with `equalize`, `grad_vol` is divided by the weights IN PLACE, like the reference
    (slice_acq_cuda_kernel.cu:1095-1107)."""
    for name, t in (("transforms", transforms), ("psf", psf), ("slices", slices), ("grad_vol", grad_vol)):
        _lib.require_cuda(name, t, slices.dtype)
    if equalize:
        _lib.require_cuda("vol", vol, slices.dtype)
        _lib.require_cuda("vol_weight", vol_weight, slices.dtype)
    else:
        vol = vol_weight = None
    vol_mask, slices_mask = _mask("vol_mask", vol_mask), _mask("slices_mask", slices_mask)
    n = transforms.shape[0]
    grad_slices = torch.zeros_like(slices) if need_slices_grad else None
    grad_tf = torch.zeros_like(transforms) if need_transforms_grad else None
    with torch.cuda.device(slices.device):
        rc = getattr(_lib.lib(), "nsv_slice_acq_adjoint_backward_" + _suffix(slices))(
            _lib.ptr(transforms), _lib.ptr(grad_vol), _lib.ptr(vol_weight), _lib.ptr(vol_mask), _lib.ptr(psf),
            _lib.ptr(slices), _lib.ptr(slices_mask), _lib.ptr(vol), _lib.ptr(grad_slices), _lib.ptr(grad_tf),
            *_dims(grad_vol.shape[-3:], psf, n, slices.shape[-2:]), _real(slices, res_slice),
            ctypes.c_int(int(interp_psf)), ctypes.c_int(int(equalize)), _lib.stream(slices.device))
    _lib.check(rc, "nsv_slice_acq_adjoint_backward")
    return [grad_slices, grad_tf]


class SliceAcqFunction(Function):
    @staticmethod
    def forward(ctx, transforms, vol, vol_mask, slices_mask, psf, slice_shape, res_slice, need_weight, interp_psf):
        if vol_mask is None:
            vol_mask = torch.empty(0, device=vol.device)
        if slices_mask is None:
            slices_mask = torch.empty(0, device=vol.device)
        outputs = forward(transforms, vol, vol_mask, slices_mask, psf, slice_shape, res_slice, need_weight, interp_psf)
        ctx.save_for_backward(transforms, vol, vol_mask, slices_mask, psf)
        ctx.interp_psf = interp_psf
        ctx.res_slice = res_slice
        ctx.need_weight = need_weight
        if need_weight:
            return outputs[0], outputs[1]
        return outputs[0]

    @staticmethod
    def backward(ctx, *args):
        if ctx.need_weight:
            assert len(args) == 2
        grad_slices = args[0]
        transforms, vol, vol_mask, slices_mask, psf = ctx.saved_tensors
        grad_vol, grad_transforms = backward(
            transforms, vol, vol_mask, psf, grad_slices.contiguous(), slices_mask, ctx.res_slice, ctx.interp_psf,
            ctx.needs_input_grad[1], ctx.needs_input_grad[0])
        return grad_transforms, grad_vol, None, None, None, None, None, None, None


class SliceAcqAdjointFunction(Function):
    @staticmethod
    def forward(ctx, transforms, psf, slices, slices_mask, vol_mask, vol_shape, res_slice, interp_psf, equalize):
        if vol_mask is None:
            vol_mask = torch.empty(0, device=slices.device)
        if slices_mask is None:
            slices_mask = torch.empty(0, device=slices.device)
        vol, vol_weight = adjoint_forward(transforms, psf, slices, slices_mask, vol_mask, vol_shape, res_slice, interp_psf, equalize)
        if equalize:
            ctx.save_for_backward(transforms, psf, slices, slices_mask, vol_mask, vol, vol_weight)
        else:
            ctx.save_for_backward(transforms, psf, slices, slices_mask, vol_mask)
        ctx.res_slice = res_slice
        ctx.interp_psf = interp_psf
        ctx.equalize = equalize
        return vol

    @staticmethod
    def backward(ctx, grad_vol):
        if ctx.equalize:
            transforms, psf, slices, slices_mask, vol_mask, vol, vol_weight = ctx.saved_tensors
        else:
            transforms, psf, slices, slices_mask, vol_mask = ctx.saved_tensors
            vol = vol_weight = None
        grad_slices, grad_transforms = adjoint_backward(
            transforms, grad_vol.contiguous(), vol_weight, vol_mask, psf, slices, slices_mask, vol, ctx.res_slice,
            ctx.interp_psf, ctx.equalize, ctx.needs_input_grad[2], ctx.needs_input_grad[0])
        return grad_transforms, None, grad_slices, None, None, None, None, None, None


def slice_acquisition(transforms, vol, vol_mask, slices_mask, psf, slice_shape, res_slice, need_weight, interp_psf):
    return SliceAcqFunction.apply(transforms, vol, vol_mask, slices_mask, psf, slice_shape, res_slice, need_weight, interp_psf)


def slice_acquisition_adjoint(transforms, psf, slices, slices_mask, vol_mask, vol_shape, res_slice, interp_psf, equalize):
    return SliceAcqAdjointFunction.apply(transforms, psf, slices, slices_mask, vol_mask, vol_shape, res_slice, interp_psf, equalize)
