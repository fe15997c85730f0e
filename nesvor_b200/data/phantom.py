"""Synthetic benchmark inputs: 3-D modified Shepp-Logan phantom and simulated slice stacks.

`phantom3d` reproduces the values of the reference's generator (tests/phantom3d.py:7-102) including
its indexing quirk (the coordinate grid has n-1 points per axis but addresses an n^3 array, which
shears the phantom).  `stack_geometry` / `simulate_stacks` follow the recipe of
tests/slice_acquisition/test_slice_acq.py:13-63: per stack a constant
rotation vector, slices spaced by `gap` along z, in-plane offset 0.5, simulated with the
slice-acquisition operator.
"""
import math
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

# A, a, b, c, x0, y0, z0, phi, theta, psi  (Toft's high-contrast head phantom)
_ELLIPSOIDS = np.array([
    [1.0, 0.6900, 0.920, 0.810, 0.00, 0.0000, 0.00, 0.0, 0.0, 0.0],
    [-0.8, 0.6624, 0.874, 0.780, 0.00, -0.0184, 0.00, 0.0, 0.0, 0.0],
    [-0.2, 0.1100, 0.310, 0.220, 0.22, 0.0000, 0.00, -18.0, 0.0, 10.0],
    [-0.2, 0.1600, 0.410, 0.280, -0.22, 0.0000, 0.00, 18.0, 0.0, 10.0],
    [0.1, 0.2100, 0.250, 0.410, 0.00, 0.3500, -0.15, 0.0, 0.0, 0.0],
    [0.1, 0.0460, 0.046, 0.050, 0.00, 0.1000, 0.25, 0.0, 0.0, 0.0],
    [0.1, 0.0460, 0.046, 0.050, 0.00, -0.1000, 0.25, 0.0, 0.0, 0.0],
    [0.1, 0.0460, 0.023, 0.050, -0.08, -0.6050, 0.00, 0.0, 0.0, 0.0],
    [0.1, 0.0230, 0.023, 0.020, 0.00, -0.6060, 0.00, 0.0, 0.0, 0.0],
    [0.1, 0.0230, 0.046, 0.020, 0.06, -0.6050, 0.00, 0.0, 0.0, 0.0],
])

STACK_ORIENTATIONS: List[Tuple[float, float, float]] = [
    (0, 0, 0), (math.pi / 2, 0, 0), (0, math.pi / 2, 0), (math.pi / 4, math.pi / 4, 0), (0, math.pi / 4, math.pi / 4),
    (math.pi / 4, 0, math.pi / 4), (math.pi / 3, math.pi / 3, 0), (0, math.pi / 3, math.pi / 3), (math.pi / 3, 0, math.pi / 3),
]


def _euler_zxz(phi: float, theta: float, psi: float) -> np.ndarray:
    cphi, sphi, cth, sth, cpsi, spsi = np.cos(phi), np.sin(phi), np.cos(theta), np.sin(theta), np.cos(psi), np.sin(psi)
    return np.array([
        [cpsi * cphi - cth * sphi * spsi, cpsi * sphi + cth * cphi * spsi, spsi * sth],
        [-spsi * cphi - cth * sphi * cpsi, -spsi * sphi + cth * cphi * cpsi, cpsi * sth],
        [sth * sphi, -sth * cphi, cth],
    ])


def phantom3d(n: int = 64) -> np.ndarray:
    """[n,n,n] float64 modified Shepp-Logan phantom, value-identical to the reference generator."""
    m = n - 1  # the quirk: m grid points per axis ...
    axis = (np.arange(m) - (n - 1) / 2) / ((n - 1) / 2)
    gx, gy, gz = np.meshgrid(axis, axis, axis)
    coord = np.vstack((gx.flatten(), gy.flatten(), gz.flatten()))
    p = np.zeros(n**3)  # ... scattered into the head of an n^3 array
    for A, a, b, c, x0, y0, z0, phi, theta, psi in _ELLIPSOIDS:
        q = np.dot(_euler_zxz(phi * np.pi / 180, theta * np.pi / 180, psi * np.pi / 180), coord)
        inside = (q[0] - x0) ** 2.0 / a**2 + (q[1] - y0) ** 2.0 / b**2 + (q[2] - z0) ** 2.0 / c**2 <= 1
        idx = np.nonzero(inside)[0]
        p[idx] = p[idx] + A
    return p.reshape((n, n, n))


def stack_geometry(n: int, res_r: float, res_s: float, gap: float, n_slice: Optional[int] = None):
    """slice side `ss` and slices per stack (test_slice_acq.py:14-19 with the volume's physical size)."""
    ss = int(math.sqrt(3) * n * res_r / res_s) + 4
    if n_slice is None:
        n_slice = int(math.sqrt(3) * n * res_r / gap) + 4
    return ss, n_slice


def stack_axisangles(orientations: Sequence[Sequence[float]], n_slice: int, gap: float, dtype=torch.float32) -> torch.Tensor:
    """[n_stacks * n_slice, 6] trans_first axis-angle rows: constant rotation, tz stepping by gap, tx = ty = 0.5."""
    rows = []
    tz = (torch.arange(n_slice, dtype=dtype) - (n_slice - 1) / 2.0) * gap
    for ang in orientations:
        a = torch.tensor([list(ang)], dtype=dtype).expand(n_slice, -1)
        t = torch.stack((torch.full_like(tz, 0.5), torch.full_like(tz, 0.5), tz), -1)
        rows.append(torch.cat((a, t), -1))
    return torch.cat(rows, 0)


def simulate_slices(n: int = 128, n_stacks: int = 3, res_r: float = 1.0, res_s: float = 1.0, gap: float = 3.0,
                    thickness: Optional[float] = None, n_slice: Optional[int] = None, device="cuda",
                    motion_deg: float = 0.0, motion_mm: float = 0.0, motion_seed: int = 1):
    """Phantom -> list of `Slice` objects, simulated with the native slice_acquisition (kernel B)
    exactly like tests/slice_acquisition/test_slice_acq.py:43-63 does.
    With `motion_*` > 0 the data are simulated at perturbed ("true") poses while the returned slices
    carry the nominal stack poses (BASELINE config 3).  Returns (slices, volume, true_axisangle)."""
    from ..image import Slice
    from ..slice_acquisition import slice_acquisition
    from ..transform import RigidTransform, mat_update_resolution
    from ..utils import get_PSF

    thickness = gap if thickness is None else thickness
    ss, n_slice = stack_geometry(n, res_r, res_s, gap, n_slice)
    volume = torch.tensor(phantom3d(n), dtype=torch.float32, device=device)[None, None]
    psf = get_PSF(res_ratio=(res_s / res_r, res_s / res_r, thickness / res_r), device=torch.device(device))
    nominal = stack_axisangles(STACK_ORIENTATIONS[:n_stacks], n_slice, gap).to(device)
    true = nominal.clone()
    if motion_deg > 0 or motion_mm > 0:
        g = torch.Generator().manual_seed(motion_seed)
        d = torch.rand(nominal.shape, generator=g) * 2 - 1
        d[:, :3] *= motion_deg * math.pi / 180.0
        d[:, 3:] *= motion_mm
        true = nominal + d.to(device)
    mat = mat_update_resolution(RigidTransform(true, trans_first=True).matrix(), 1, res_r).contiguous()
    images = slice_acquisition(mat, volume, None, None, psf, (ss, ss), res_s / res_r, False, False)
    slices = []
    nominal_t = RigidTransform(nominal, trans_first=True)
    for i in range(images.shape[0]):
        img = images[i]
        mask = img > 0
        if not mask.any():
            continue  # empty slices are dropped, as svort/inference.py:557-559 does
        slices.append(Slice(img, mask, nominal_t[i], res_s, res_s, thickness, stack_idx=i // n_slice, slice_idx=i % n_slice))
    return slices, volume, true
