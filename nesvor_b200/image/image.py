"""Minimal Image / Slice / Volume containers feeding the hot path.

Only what `train()` / `sample_*()` touch is mirrored from nesvor/image/image.py: constructor
signature (:17-42), `shape_xyz` / `resolution_xyz` (:57-66), `xyz_masked`, `xyz_masked_untransformed`,
`v_masked` (:80-90), `clone` (:112-120), `Volume.sample_points` / `resample` / `xyz_masked`
(:123-183).  NIfTI I/O (nibabel, absent in this image) is out of scope.
"""
from __future__ import annotations

from typing import Dict, Optional, Union

import torch
import torch.nn.functional as F

from ..transform import RigidTransform, transform_points
from ..utils.misc import meshgrid


class Image(object):
    def __init__(self, image: torch.Tensor, mask: Optional[torch.Tensor] = None,
                 transformation: Optional[RigidTransform] = None, resolution_x: Union[float, torch.Tensor] = 1.0,
                 resolution_y: Union[float, torch.Tensor] = 1.0, resolution_z: Union[float, torch.Tensor] = 1.0) -> None:
        assert image.ndim == 3
        self.image = image
        self.mask = torch.ones_like(image, dtype=torch.bool) if mask is None else mask
        if transformation is None:
            transformation = RigidTransform(torch.zeros((1, 6), dtype=torch.float32, device=image.device))
        self.transformation = transformation
        self.resolution_x, self.resolution_y, self.resolution_z = resolution_x, resolution_y, resolution_z

    def _clone_image(self, zero: bool = False) -> Dict:
        return {
            "image": torch.zeros_like(self.image) if zero else self.image.clone(),
            "mask": torch.zeros_like(self.mask) if zero else self.mask.clone(),
            "transformation": self.transformation.clone(),
            "resolution_x": float(self.resolution_x),
            "resolution_y": float(self.resolution_y),
            "resolution_z": float(self.resolution_z),
        }

    @property
    def shape_xyz(self) -> torch.Tensor:
        return torch.tensor(self.image.shape[::-1], device=self.image.device)

    @property
    def resolution_xyz(self) -> torch.Tensor:
        return torch.tensor([self.resolution_x, self.resolution_y, self.resolution_z], device=self.image.device)

    @property
    def xyz_masked(self) -> torch.Tensor:
        return transform_points(self.transformation, self.xyz_masked_untransformed)

    @property
    def xyz_masked_untransformed(self) -> torch.Tensor:
        kji = torch.flip(torch.nonzero(self.mask), (-1,))
        return (kji - (self.shape_xyz - 1) / 2) * self.resolution_xyz

    @property
    def v_masked(self) -> torch.Tensor:
        return self.image[self.mask]


class Slice(Image):
    def __init__(self, image, mask=None, transformation=None, resolution_x=1.0, resolution_y=1.0, resolution_z=1.0,
                 stack_idx: Optional[int] = None, slice_idx: Optional[int] = None) -> None:
        super().__init__(image, mask, transformation, resolution_x, resolution_y, resolution_z)
        self.stack_idx, self.slice_idx = stack_idx, slice_idx

    def clone(self, zero: bool = False) -> "Slice":
        return Slice(stack_idx=self.stack_idx, slice_idx=self.slice_idx, **self._clone_image(zero))


class Volume(Image):
    def clone(self, zero: bool = False) -> "Volume":
        return Volume(**self._clone_image(zero))

    def sample_points(self, xyz: torch.Tensor) -> torch.Tensor:
        """Trilinear sample of the volume at world points (image.py:123-133)."""
        shape = xyz.shape[:-1]
        xyz = transform_points(self.transformation.inv(), xyz.view(-1, 3))
        xyz = xyz / ((self.shape_xyz - 1) * self.resolution_xyz / 2)
        return F.grid_sample(self.image[None, None], xyz.view(1, 1, 1, -1, 3), align_corners=True).view(shape)

    def resample(self, resolution_new, transformation_new: Optional[RigidTransform]) -> "Volume":
        """New axis-aligned grid at `resolution_new` covering the mask (image.py:135-181)."""
        if transformation_new is None:
            transformation_new = self.transformation
        R = transformation_new.matrix()[0, :3, :3]
        dtype, device = R.dtype, R.device
        if isinstance(resolution_new, (float, int)) or getattr(resolution_new, "numel", lambda: 3)() == 1:
            resolution_new = torch.tensor([float(resolution_new)] * 3, dtype=dtype, device=device)
        xyz = self.xyz_masked
        xyz = torch.matmul(torch.inverse(R), xyz.view(-1, 3, 1))[..., 0]
        xyz_min = xyz.amin(0) - resolution_new * 10
        xyz_max = xyz.amax(0) + resolution_new * 10
        shape_xyz = ((xyz_max - xyz_min) / resolution_new).ceil().long()
        mat = torch.zeros((1, 3, 4), dtype=dtype, device=device)
        mat[0, :, :3] = R
        mat[0, :, 3] = xyz_min + (shape_xyz - 1) / 2 * resolution_new
        xyz = meshgrid(shape_xyz, resolution_new, xyz_min, device, True)
        xyz = torch.matmul(R, xyz[..., None])[..., 0]
        v = self.sample_points(xyz)
        return Volume(v, v > 0, RigidTransform(mat, trans_first=True), resolution_new[0].item(),
                      resolution_new[1].item(), resolution_new[2].item())
