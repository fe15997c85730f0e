"""NIfTI affine <-> per-slice rigid transforms (the geometry contract of nesvor/image/image_utils.py:8-85, restated).

A NIfTI affine A = [M | o] maps voxel indices (i, j, k) to scanner millimetres.  The reconstruction path wants, for every
slice k of a stack, a rigid transform in the `trans_first` convention  x_world = Q (x_slice + t_k),  x_slice in millimetres
from the centre of the slice.  With the voxel spacings s = (s_x, s_y, s_z):

    Q   = M diag(1/s)                                   (direction cosines)
    t_k = Q^-1 o + ((w-1)/2 s_x, (h-1)/2 s_y, k s_z)    (voxel (0,0,k) seen from the slice centre)

and the inverse for a single volume-centred transform (`transformation2affine`).  Left-handed affines (det M < 0) are
made right-handed the way the reference does: the image is mirrored along x, and the first column of Q and the x
translation change sign.
"""
from typing import Tuple

import numpy as np
import torch

from ..transform import RigidTransform

_TOL = 1e-3  # image_utils.py:17,21


def compare_resolution_affine(r1, a1, r2, a2, s1, s2) -> bool:
    """True when two images share shape, voxel spacings and affine (1e-3 absolute), as a stack and its mask must."""
    if s1 != s2:
        return False
    for u, v in ((np.asarray(r1, np.float64), np.asarray(r2, np.float64)), (np.asarray(a1, np.float64), np.asarray(a2, np.float64))):
        if u.shape != v.shape or np.abs(u - v).max() > _TOL:
            return False
    return True


def affine2transformation(volume: torch.Tensor, mask: torch.Tensor, resolutions, affine) -> Tuple[torch.Tensor, torch.Tensor, RigidTransform]:
    """[d, h, w] image + NIfTI affine -> (image, mask, RigidTransform holding one [Q | t_k] per slice k)."""
    d, h, w = volume.shape
    s = np.asarray(resolutions, np.float64).reshape(3)
    A = np.asarray(affine, np.float64)
    Q = A[:3, :3] / s  # scales column j by 1 / s_j
    left_handed = np.linalg.det(A[:3, :3]) < 0
    t0 = np.linalg.solve(Q, A[:3, 3]) + np.array([(w - 1) / 2 * s[0], (h - 1) / 2 * s[1], 0.0])
    t = np.tile(t0, (d, 1))
    t[:, 2] += np.arange(d) * s[2]
    Qs = np.tile(Q, (d, 1, 1))
    if left_handed:
        volume, mask = torch.flip(volume, (-1,)), torch.flip(mask, (-1,))
        t[:, 0] = -t[:, 0]
        Qs[:, :, 0] = -Qs[:, :, 0]
    mat = torch.tensor(np.concatenate([Qs, t[:, :, None]], -1), dtype=torch.float32, device=volume.device)
    return volume, mask, RigidTransform(mat, trans_first=True)


def transformation2affine(volume: torch.Tensor, transformation: RigidTransform, resolution_x: float, resolution_y: float,
                          resolution_z: float) -> np.ndarray:
    """One volume-centred rigid transform + spacings -> NIfTI affine (voxel (0,0,0) sits half an extent before the centre)."""
    mat = transformation.matrix(trans_first=True).detach().cpu().numpy().astype(np.float64)
    if mat.shape[0] != 1:
        raise ValueError("transformation2affine expects a single transform")
    d, h, w = volume.shape
    s = np.array([resolution_x, resolution_y, resolution_z], np.float64)
    Q, t = mat[0, :, :3], mat[0, :, 3]
    affine = np.eye(4)
    affine[:3, :3] = Q * s
    affine[:3, 3] = Q @ (t - (np.array([w, h, d], np.float64) - 1) / 2 * s)
    return affine
