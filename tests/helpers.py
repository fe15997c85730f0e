"""Shared builders for tests: seeded geometry for kernel B, INR configs, error metrics."""
import math

import numpy as np
import torch


def rel_l2(a, b) -> float:
    a = torch.as_tensor(a, dtype=torch.float64).flatten()
    b = torch.as_tensor(b, dtype=torch.float64).flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def rotvec_to_mat(v: np.ndarray) -> np.ndarray:
    from scipy.spatial.transform import Rotation

    return Rotation.from_rotvec(v).as_matrix()


def gaussian_psf(res_ratio=(1.5, 1.5, 3.0), dtype=np.float32) -> np.ndarray:
    """get_PSF of the product package, as numpy (its values are pinned by tests/golden/psf_*.npy)."""
    from nesvor_b200.utils.psf import get_PSF

    return get_PSF(res_ratio=res_ratio).numpy().astype(dtype)


def slice_acq_case(seed=0, dtype=np.float32, D=20, H=22, W=24, n=5, h=18, w=17, masks=False, res_ratio=(1.5, 1.5, 3.0)):
    rng = np.random.default_rng(seed)
    vol = rng.random((1, 1, D, H, W)).astype(dtype)
    psf = gaussian_psf(res_ratio, dtype)
    tf = np.zeros((n, 3, 4), dtype)
    for i in range(n):
        tf[i, :, :3] = rotvec_to_mat(rng.normal(size=3) * 0.7)
        tf[i, :, 3] = rng.normal(size=3) * 2.0
    vol_mask = (rng.random((1, 1, D, H, W)) > 0.2) if masks else None
    slices_mask = (rng.random((n, 1, h, w)) > 0.2) if masks else None
    slices = rng.random((n, 1, h, w)).astype(dtype)
    grad_slices = rng.normal(size=(n, 1, h, w)).astype(dtype)
    grad_slices[0, 0, :3] = 0  # exercises the gs == 0 early-out (Q2)
    grad_vol = rng.normal(size=(1, 1, D, H, W)).astype(dtype)
    return dict(vol=vol, psf=psf, transforms=tf, vol_mask=vol_mask, slices_mask=slices_mask, slices=slices,
                grad_slices=grad_slices, grad_vol=grad_vol, slice_shape=(h, w), vol_shape=(D, H, W), res_slice=1.5)


def cuda(x, device="cuda"):
    if x is None:
        return None
    return torch.as_tensor(x).to(device).contiguous()


REF_AXISANGLES = [  # the reference's 11 hand-picked vectors, tests/__init__.py:24-36
    [0, 0, 0, 0, 0, 0],
    [np.pi / 2, 0, 0, 1, 2, 3],
    [0, -np.pi / 2, 0, -1.1, -10, 100.5],
    [0, 0, np.pi - 0.01, 2, 1, 10.5],
    [0, -np.pi + 0.01, 0, 2, 1, 10.5],
    [0.1, 0.1, 0.1, 0.1, 0.1, 0.1],
    [-0.1, 0, -0.4, 0.1, 0.5, 0.1],
    [-0.2, 0.2, -0.1, -100, 200, -159],
    [-0.12, -0.01, 0.1, -100, 200, -159],
    [np.pi / 4, np.pi / 4, np.pi / 4, 0.1, 0.1, 0.1],
    [np.pi / 3, -np.pi / 4, np.pi / 5, 100, 200, -300],
]


def scipy_axisangle2mat(ax: np.ndarray) -> np.ndarray:
    from scipy.spatial.transform import Rotation

    mat = Rotation.from_rotvec(ax[:, :3].astype(np.float64)).as_matrix()
    return np.concatenate([mat, ax[:, 3:, None]], -1).astype(ax.dtype)
