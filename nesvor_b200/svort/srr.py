"""Super-resolution reconstruction on the slice-acquisition operator (kernel B): the step *before* the INR path
(initial volume, SVoRT's inner loop).

Host-side mirror of nesvor/svort/srr.py: `CG` (:12-34), `PSFreconstruction` (:37-48) and `SRR` (:51-160) with the same
names, argument order and `params` dictionary keys ("psf", "slice_shape", "volume_shape", "res_s", "res_r",
"interp_psf").  All the arithmetic that touches slices or volumes runs in nsv_slice_acq_* (A, A^T) -- the solver around
them is a handful of dot products and axpys on device tensors.
"""
from typing import Callable, Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..slice_acquisition import slice_acquisition, slice_acquisition_adjoint
from ..transform import axisangle2mat


def _dot(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    return torch.dot(x.flatten(), y.flatten())


def CG(A: Callable[[torch.Tensor], torch.Tensor], b: torch.Tensor, x0: Optional[torch.Tensor], n_iter: int, tol: float = 0.0):
    """Conjugate gradients for A x = b (A symmetric positive semi-definite, given as a callable).  Same stopping rules
    as the reference: exactly `n_iter` updates of x unless the squared residual drops to `tol` first; x0 = None starts
    from zero without evaluating A(0)."""
    if x0 is None:
        x, r = 0, b
    else:
        x, r = x0, b - A(x0)
    p = r
    rr = _dot(r, r)
    for i in range(1, n_iter + 1):
        Ap = A(p)
        alpha = rr / _dot(p, Ap)
        x = x + alpha * p
        if i == n_iter:
            break
        r = r - alpha * Ap
        rr_new = _dot(r, r)
        if rr_new <= tol:
            break
        p = r + (rr_new / rr) * p
        rr = rr_new
    return x


def PSFreconstruction(transforms, slices, slices_mask, vol_mask, params: Dict):
    """Weight-normalised scatter of the slices into the volume (A^T with `equalize`): the initial volume."""
    return slice_acquisition_adjoint(transforms, params["psf"], slices, slices_mask, vol_mask, params["volume_shape"],
                                     params["res_s"] / params["res_r"], params["interp_psf"], True)


class SRR(nn.Module):
    """min_x |p^(1/2) (A x - y)|^2 (+ mu |x - z|^2): `n_iter` CG iterations on the normal equations, or `n_iter` steps
    of gradient descent with step `alpha` and the edge-preserving regulariser of weight beta * delta^2."""

    def __init__(self, n_iter: int = 10, use_CG: bool = False, alpha: float = 0.5, beta: float = 0.02, delta: float = 0.1, tol: float = 0.0):
        super().__init__()
        self.n_iter, self.use_CG, self.alpha, self.delta, self.tol = n_iter, use_CG, alpha, delta, tol
        self.beta = beta * delta * delta

    def forward(self, theta, slices, volume, params: Dict, p=None, mu=0, z=None, vol_mask=None, slices_mask=None):
        transforms = axisangle2mat(theta) if theta.ndim == 2 else theta

        def A(x):
            return self.A(transforms, x, vol_mask, slices_mask, params)

        def At(y):
            return self.At(transforms, y, slices_mask, vol_mask, params)

        if self.use_CG:
            b = At(slices * p if p is not None else slices)
            if mu and z is not None:
                b = b + mu * z
            x = CG(lambda x: self.AtA(transforms, x, vol_mask, slices_mask, p, params, mu, z), b, volume, self.n_iter, self.tol)
        else:
            x = volume
            for _ in range(self.n_iter):
                err = A(x) - slices
                if p is not None:
                    err = p * err
                g = At(err)
                if self.beta:
                    g.add_(self.dR(x, self.delta), alpha=self.beta)
                x.add_(g, alpha=-self.alpha)  # in place, like the reference: the caller's volume is updated
        return F.relu(x, True)

    def A(self, transforms, x, vol_mask, slices_mask, params):
        return slice_acquisition(transforms, x, vol_mask, slices_mask, params["psf"], params["slice_shape"],
                                 params["res_s"] / params["res_r"], False, params["interp_psf"])

    def At(self, transforms, x, slices_mask, vol_mask, params):
        return slice_acquisition_adjoint(transforms, params["psf"], x, slices_mask, vol_mask, params["volume_shape"],
                                         params["res_s"] / params["res_r"], params["interp_psf"], False)

    def AtA(self, transforms, x, vol_mask, slices_mask, p, params, mu, z):
        s = self.A(transforms, x, vol_mask, slices_mask, params)
        if p is not None:
            s = s * p
        vol = self.At(transforms, s, slices_mask, vol_mask, params)
        if mu and z is not None:
            vol = vol + mu * x
        return vol

    @staticmethod
    def dR(v: torch.Tensor, delta: float) -> torch.Tensor:
        """Edge-preserving prior term of the gradient-descent branch, as the reference evaluates it (srr.py:139-160):
        for every interior voxel, summed over its 26 neighbours n, with d = v - v_n and s = d / (|n|^2 delta^2):
        s / sqrt(1 + d s).  Border voxels get zero."""
        g = torch.zeros_like(v)
        D, H, W = v.shape[-3:]
        c = v[..., 1 : D - 1, 1 : H - 1, 1 : W - 1]
        gi = g[..., 1 : D - 1, 1 : H - 1, 1 : W - 1]
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    if dx == dy == dz == 0:
                        continue
                    d = c - v[..., 1 + dz : D - 1 + dz, 1 + dy : H - 1 + dy, 1 + dx : W - 1 + dx]
                    s = d * (1.0 / (dx * dx + dy * dy + dz * dz) / (delta * delta))
                    gi += s / torch.sqrt(1 + d * s)
        return g
