"""TEST INFRASTRUCTURE ONLY.  ctypes/numpy front-end for the two CPU libraries of oracle/:

* ``Oracle()``  -> oracle/_build/libnsv_oracle.so  (the C restatement, always buildable)
* ``Reference()`` -> oracle/_ref/libnesvor_ref_cpu.so (the reference's own kernels on CPU; None when
  neither /root/reference nor a prebuilt copy exists)

Both expose the argument lists of the reference's pybind modules
(nesvor/slice_acquisition/slice_acq_cuda.cpp:61-153, nesvor/transform/transform_convert_cuda.cpp:27-61)
on numpy arrays; masks are ``None`` or bool arrays; outputs are fresh arrays.
"""
import ctypes
import os

import numpy as np

from . import build as _build

_c_int = ctypes.c_int
_ptr = ctypes.c_void_p


def _p(a):
    return None if a is None else a.ctypes.data_as(_ptr)


class _KernelLib:
    def __init__(self, path, prefix, long_equalize):
        self.path = path
        self.lib = ctypes.CDLL(path)
        self.prefix = prefix
        self.long_equalize = long_equalize

    # ------------------------------------------------------------------ helpers
    def _fn(self, name, dtype):
        suffix = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[np.dtype(dtype)]
        f = getattr(self.lib, f"{self.prefix}{name}_{suffix}")
        f.restype = None
        return f

    @staticmethod
    def _real(dtype, v):
        return ctypes.c_float(v) if np.dtype(dtype) == np.float32 else ctypes.c_double(v)

    @staticmethod
    def _prep(dtype, *arrs):
        out = []
        for a in arrs:
            out.append(None if a is None else np.ascontiguousarray(a, dtype=dtype))
        return out

    @staticmethod
    def _mask(m):
        if m is None or m.size == 0:
            return None
        return np.ascontiguousarray(m, dtype=np.bool_)

    @staticmethod
    def _dims(vol_shape, psf, n, slice_shape):
        D, H, W = [int(v) for v in vol_shape]
        d_p, h_p, w_p = [int(v) for v in psf.shape]
        h, w = [int(v) for v in slice_shape]
        return [_c_int(v) for v in (D, H, W, d_p, h_p, w_p, int(n), h, w)]

    # ------------------------------------------------------------------ slice acquisition
    def forward(self, transforms, vol, vol_mask, slices_mask, psf, slice_shape, res_slice, need_weight, interp_psf):
        dt = vol.dtype
        transforms, vol, psf = self._prep(dt, transforms, vol, psf)
        vol_mask, slices_mask = self._mask(vol_mask), self._mask(slices_mask)
        n = transforms.shape[0]
        slices = np.empty((n, 1, slice_shape[0], slice_shape[1]), dt)
        weight = np.empty_like(slices) if need_weight else None
        self._fn("slice_acq_forward", dt)(
            _p(transforms), _p(vol), _p(vol_mask), _p(slices_mask), _p(psf), _p(slices), _p(weight),
            *self._dims(vol.shape[-3:], psf, n, slice_shape), self._real(dt, res_slice), _c_int(int(interp_psf)))
        return [slices, weight] if need_weight else [slices]

    def backward(self, transforms, vol, vol_mask, psf, grad_slices, slices_mask, res_slice, interp_psf,
                 need_vol_grad, need_transforms_grad):
        dt = vol.dtype
        transforms, vol, psf, grad_slices = self._prep(dt, transforms, vol, psf, grad_slices)
        vol_mask, slices_mask = self._mask(vol_mask), self._mask(slices_mask)
        n = transforms.shape[0]
        grad_vol = np.empty_like(vol) if need_vol_grad else None
        grad_tf = np.empty_like(transforms) if need_transforms_grad else None
        self._fn("slice_acq_backward", dt)(
            _p(transforms), _p(vol), _p(vol_mask), _p(psf), _p(grad_slices), _p(slices_mask), _p(grad_vol), _p(grad_tf),
            *self._dims(vol.shape[-3:], psf, n, grad_slices.shape[-2:]), self._real(dt, res_slice), _c_int(int(interp_psf)))
        return [grad_vol, grad_tf]

    def adjoint_forward(self, transforms, psf, slices, slices_mask, vol_mask, vol_shape, res_slice, interp_psf, equalize):
        dt = slices.dtype
        transforms, psf, slices = self._prep(dt, transforms, psf, slices)
        vol_mask, slices_mask = self._mask(vol_mask), self._mask(slices_mask)
        n = transforms.shape[0]
        vol = np.empty((1, 1) + tuple(int(v) for v in vol_shape), dt)
        vol_weight = np.empty_like(vol) if equalize else None
        self._fn("slice_acq_adjoint_forward", dt)(
            _p(transforms), _p(psf), _p(slices), _p(slices_mask), _p(vol_mask), _p(vol), _p(vol_weight),
            *self._dims(vol_shape, psf, n, slices.shape[-2:]), self._real(dt, res_slice), _c_int(int(interp_psf)),
            _c_int(int(equalize)))
        return [vol, vol_weight]

    def adjoint_backward(self, transforms, grad_vol, vol_weight, vol_mask, psf, slices, slices_mask, vol, res_slice,
                         interp_psf, equalize, need_slices_grad, need_transforms_grad):
        """NB: like the reference, ``grad_vol`` is modified in place when ``equalize`` (a copy is
        made here only if the caller's array is not already contiguous and of the right dtype)."""
        dt = slices.dtype
        transforms, psf, slices = self._prep(dt, transforms, psf, slices)
        grad_vol = np.ascontiguousarray(grad_vol, dtype=dt)
        if equalize:
            vol_weight, vol = self._prep(dt, vol_weight, vol)
        else:
            vol_weight = vol = None
        vol_mask, slices_mask = self._mask(vol_mask), self._mask(slices_mask)
        n = transforms.shape[0]
        grad_slices = np.empty_like(slices) if need_slices_grad else None
        grad_tf = np.empty_like(transforms) if need_transforms_grad else None
        self._fn("slice_acq_adjoint_backward", dt)(
            _p(transforms), _p(grad_vol), _p(vol_weight), _p(vol_mask), _p(psf), _p(slices), _p(slices_mask), _p(vol),
            _p(grad_slices), _p(grad_tf), *self._dims(grad_vol.shape[-3:], psf, n, slices.shape[-2:]),
            self._real(dt, res_slice), _c_int(int(interp_psf)), _c_int(int(equalize)))
        return [grad_slices, grad_tf]

    def equalize(self, vol, vol_weight, is_grad):
        dt = vol.dtype
        assert vol.flags.c_contiguous
        (vol_weight,) = self._prep(dt, vol_weight)
        cnt = ctypes.c_long(vol.size) if self.long_equalize else _c_int(vol.size)
        self._fn("equalize", dt)(_p(vol), _p(vol_weight), _c_int(int(is_grad)), cnt)
        return vol

    # ------------------------------------------------------------------ pose converters
    def axisangle2mat_forward(self, axisangle):
        (ax,) = self._prep(axisangle.dtype, axisangle)
        mat = np.empty((ax.shape[0], 3, 4), ax.dtype)
        self._fn("axisangle2mat_forward", ax.dtype)(_p(ax), _p(mat), _c_int(ax.shape[0]))
        return [mat]

    def axisangle2mat_backward(self, grad_mat, axisangle):
        grad_mat, ax = self._prep(axisangle.dtype, grad_mat, axisangle)
        g = np.empty_like(ax)
        self._fn("axisangle2mat_backward", ax.dtype)(_p(grad_mat), _p(ax), _p(g), _c_int(ax.shape[0]))
        return [g]

    def mat2axisangle_forward(self, mat):
        (mat,) = self._prep(mat.dtype, mat)
        ax = np.empty((mat.shape[0], 6), mat.dtype)
        self._fn("mat2axisangle_forward", mat.dtype)(_p(mat), _p(ax), _c_int(mat.shape[0]))
        return [ax]

    def mat2axisangle_backward(self, mat, grad_axisangle):
        mat, ga = self._prep(mat.dtype, mat, grad_axisangle)
        g = np.empty_like(mat)
        self._fn("mat2axisangle_backward", mat.dtype)(_p(mat), _p(ga), _p(g), _c_int(mat.shape[0]))
        return [g]


_ORACLE = None
_REFERENCE = False


def Oracle():
    """The C restatement (built on demand with gcc)."""
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = _KernelLib(_build.build_oracle(), "nsv_oracle_", long_equalize=True)
    return _ORACLE


def Reference():
    """The reference's own kernels compiled for CPU, or None when unavailable."""
    global _REFERENCE
    if _REFERENCE is False:
        path = _build.build_ref()
        _REFERENCE = _KernelLib(path, "nsv_ref_", long_equalize=False) if path else None
    return _REFERENCE


def set_threads(n):
    """OpenMP thread count for both libraries (1 = deterministic summation order)."""
    os.environ["OMP_NUM_THREADS"] = str(n)
    for lib in (Oracle(), Reference()):
        if lib is None:
            continue
        try:
            f = lib.lib.omp_set_num_threads
        except AttributeError:
            continue
        f(_c_int(int(n)))
