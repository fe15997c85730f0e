"""PSF constants and the discrete Gaussian PSF used by slice_acquisition.

Mirrors nesvor/utils/psf.py:5-65 (GAUSSIAN_FWHM, SINC_FWHM, resolution2sigma, get_PSF).
"""
from math import log, sqrt
from typing import Optional, Sequence

import torch

GAUSSIAN_FWHM = 1 / (2 * sqrt(2 * log(2)))
SINC_FWHM = 1.206709128803223 * GAUSSIAN_FWHM


def resolution2sigma(rx, ry=None, rz=None, /, isotropic=False):
    """Resolution (mm) -> Gaussian sigma: in-plane sinc main lobe, through-plane Gaussian profile."""
    fx, fy, fz = (GAUSSIAN_FWHM,) * 3 if isotropic else (SINC_FWHM, SINC_FWHM, GAUSSIAN_FWHM)
    assert (ry is None) == (rz is None)
    if ry is not None:
        return fx * rx, fy * ry, fz * rz
    if isinstance(rx, (float, int)):
        return fx * rx if isotropic else (fx * rx, fy * rx, fz * rx)
    if isinstance(rx, torch.Tensor):
        if isotropic:
            return fx * rx
        assert rx.shape[-1] == 3
        return rx * torch.tensor([fx, fy, fz], dtype=rx.dtype, device=rx.device)
    if isinstance(rx, (list, tuple)):
        assert len(rx) == 3
        return resolution2sigma(rx[0], rx[1], rx[2], isotropic=isotropic)
    raise Exception(str(type(rx)))


def get_PSF(r_max: Optional[int] = None, res_ratio: Sequence[float] = (1, 1, 3), threshold: float = 1e-3,
            device=torch.device("cpu")) -> torch.Tensor:
    """[d_p, h_p, w_p] Gaussian on the reconstruction-voxel grid, values below `threshold` zeroed,
    cropped to its support and normalised to sum 1."""
    sx, sy, sz = resolution2sigma(tuple(res_ratio), isotropic=False)
    if r_max is None:
        r_max = max(max(int(2 * r + 1) for r in (sx, sy, sz)), 4)
    ax = torch.linspace(-r_max, r_max, 2 * r_max + 1, dtype=torch.float32, device=device)
    gz, gy, gx = torch.meshgrid(ax, ax, ax, indexing="ij")
    psf = torch.exp(-0.5 * (gx**2 / sx**2 + gy**2 / sy**2 + gz**2 / sz**2))
    psf[psf.abs() < threshold] = 0
    side = 2 * r_max + 1
    cx = int(torch.nonzero(psf.sum((0, 1)) > 0)[0, 0].item())
    cy = int(torch.nonzero(psf.sum((0, 2)) > 0)[0, 0].item())
    cz = int(torch.nonzero(psf.sum((1, 2)) > 0)[0, 0].item())
    psf = psf[cz : side - cz, cy : side - cy, cx : side - cx].contiguous()
    return psf / psf.sum()
