"""axis-angle <-> matrix converters on the native library.

Mirrors nesvor/transform/transform_convert.py:20-57 (Axisangle2MatFunction, Mat2AxisangleFunction,
axisangle2mat, mat2axisangle): same names, argument meaning and error behaviour (CUDA, contiguous
tensors only); float32 and float64 are dispatched like the reference's AT_DISPATCH_FLOATING_TYPES.
Launches go to torch's current stream (the reference uses the legacy default stream).
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
    raise RuntimeError(f"nesvor_b200 pose converters support float32/float64, got {t.dtype}")


def _call(name: str, n: int, *tensors: torch.Tensor) -> None:
    fn = getattr(_lib.lib(), f"{name}_{_suffix(tensors[0])}")
    with torch.cuda.device(tensors[0].device):
        rc = fn(*[_lib.ptr(t) for t in tensors], ctypes.c_int(n), _lib.stream(tensors[0].device))
    _lib.check(rc, name)


def axisangle2mat_forward(axisangle: torch.Tensor):
    _lib.require_cuda("axisangle", axisangle)
    mat = torch.empty((axisangle.shape[0], 3, 4), dtype=axisangle.dtype, device=axisangle.device)
    _call("nsv_axisangle2mat_fwd", axisangle.shape[0], axisangle, mat)
    return [mat]


def axisangle2mat_backward(grad_mat: torch.Tensor, axisangle: torch.Tensor):
    _lib.require_cuda("axisangle", axisangle)
    _lib.require_cuda("grad_mat", grad_mat, axisangle.dtype)
    grad = torch.empty_like(axisangle)
    _call("nsv_axisangle2mat_bwd", axisangle.shape[0], grad_mat, axisangle, grad)
    return [grad]


def mat2axisangle_forward(mat: torch.Tensor):
    _lib.require_cuda("mat", mat)
    ax = torch.empty((mat.shape[0], 6), dtype=mat.dtype, device=mat.device)
    _call("nsv_mat2axisangle_fwd", mat.shape[0], mat, ax)
    return [ax]


def mat2axisangle_backward(mat: torch.Tensor, grad_axisangle: torch.Tensor):
    _lib.require_cuda("mat", mat)
    _lib.require_cuda("grad_axisangle", grad_axisangle, mat.dtype)
    grad = torch.empty_like(mat)
    _call("nsv_mat2axisangle_bwd", mat.shape[0], mat, grad_axisangle, grad)
    return [grad]


class Axisangle2MatFunction(Function):
    @staticmethod
    def forward(ctx, axisangle):
        ctx.save_for_backward(axisangle)
        return axisangle2mat_forward(axisangle)[0]

    @staticmethod
    def backward(ctx, grad_mat):
        (axisangle,) = ctx.saved_tensors
        return axisangle2mat_backward(grad_mat.contiguous(), axisangle)[0]


class Mat2AxisangleFunction(Function):
    @staticmethod
    def forward(ctx, mat):
        ctx.save_for_backward(mat)
        return mat2axisangle_forward(mat)[0]

    @staticmethod
    def backward(ctx, grad_axisangle):
        (mat,) = ctx.saved_tensors
        return mat2axisangle_backward(mat, grad_axisangle.contiguous())[0]


def axisangle2mat(axisangle: torch.Tensor) -> torch.Tensor:
    return Axisangle2MatFunction.apply(axisangle)


def mat2axisangle(mat: torch.Tensor) -> torch.Tensor:
    return Mat2AxisangleFunction.apply(mat)
