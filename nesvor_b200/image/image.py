"""Image / Slice / Volume / Stack containers and their NIfTI I/O: the data formats either side of the hot path.

Mirrors nesvor/image/image.py: constructor signature (:17-42), `shape_xyz` / `resolution_xyz` (:57-66), `save` (:68-82),
`xyz_masked`, `xyz_masked_untransformed`, `v_masked` (:84-96), `rescale` (:98-100), `clone` (:112-120),
`Volume.sample_points` / `resample` (:123-183), `Stack` (:186-250), `save_nii_volume` / `load_nii_volume` (:253-296),
`save_slices` / `load_slices` (:299-330), `load_stack` (:333-365), `load_volume` (:368-400).  The reference goes through
nibabel, which is absent in this image: the few calls it makes are restated in `nifti.py`.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

from ..transform import RigidTransform, transform_points
from ..utils.misc import meshgrid
from .affine import affine2transformation, compare_resolution_affine, transformation2affine
from .nifti import read_nifti, write_nifti


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

    def save(self, path: str, masked: bool = True) -> None:
        """NIfTI file of the (masked) image with the affine of its pose and spacings (image.py:68-82)."""
        affine = transformation2affine(self.image, self.transformation, float(self.resolution_x), float(self.resolution_y),
                                       float(self.resolution_z))
        out = self.image * self.mask.to(self.image.dtype) if masked else self.image
        save_nii_volume(path, out, affine)

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

    def rescale(self, intensity_mean: Union[float, torch.Tensor]) -> None:
        self.image *= intensity_mean / self.image[self.mask].mean()


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


class Stack(object):
    """A stack of 2-D slices [n, 1, h, w] with one rigid transform per slice (image.py:186-250)."""

    def __init__(self, slices: torch.Tensor, mask: Optional[torch.Tensor] = None, transformation: Optional[RigidTransform] = None,
                 score: float = 0.0, resolution_x: float = 1.0, resolution_y: float = 1.0, thickness: float = 1.0, gap: float = 1.0) -> None:
        self.slices = slices
        self.mask = torch.ones_like(slices, dtype=torch.bool) if mask is None else mask
        if transformation is None:  # slices spaced by `gap` along z, centred on the stack
            n = slices.shape[0]
            t = torch.zeros((n, 6), dtype=torch.float32, device=slices.device)
            t[:, -1] = (torch.arange(n, dtype=torch.float32, device=slices.device) - n / 2) * gap
            transformation = RigidTransform(t)
        self.transformation = transformation
        self.score = torch.ones(slices.shape[0], dtype=torch.float32, device=slices.device) if score is None else score
        self.resolution_x, self.resolution_y, self.thickness, self.gap = resolution_x, resolution_y, thickness, gap

    def __len__(self) -> int:
        return self.slices.shape[0]

    def __getitem__(self, idx):
        assert self.slices.ndim == 4
        slices, masks, transformation = self.slices[idx], self.mask[idx], self.transformation[idx]
        if slices.ndim < self.slices.ndim:
            return Slice(slices, masks, transformation, self.resolution_x, self.resolution_y, self.thickness)
        return [Slice(slices[i], masks[i], transformation[i], self.resolution_x, self.resolution_y, self.thickness)
                for i in range(len(transformation))]


def save_nii_volume(path: str, volume: Union[torch.Tensor, np.ndarray], affine: Optional[Union[torch.Tensor, np.ndarray]]) -> None:
    """[d, h, w] (or [d, 1, h, w]) -> NIfTI file with x fastest; qform "aligned" + sform "scanner", mm (image.py:253-272)."""
    assert len(volume.shape) == 3 or (len(volume.shape) == 4 and volume.shape[1] == 1)
    if len(volume.shape) == 4:
        volume = volume.squeeze(1)
    if isinstance(volume, torch.Tensor):
        volume = volume.detach().cpu().numpy()
    if isinstance(affine, torch.Tensor):
        affine = affine.detach().cpu().numpy()
    write_nifti(path, np.transpose(volume, (2, 1, 0)), affine)


def load_nii_volume(path: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """NIfTI file -> (float32 [d, h, w], spacings (x, y, z), 4x4 affine) (image.py:275-296)."""
    data, hdr = read_nifti(path)
    dim = hdr["dim"]
    if not (dim[0] == 3 or (dim[0] > 3 and all(int(d) == 1 for d in dim[4 : 1 + dim[0]]))):
        raise AssertionError("Expect a 3D volume but the input is %dD" % dim[0])
    volume = data.astype(np.float32)
    while volume.ndim > 3:
        volume = volume.squeeze(-1)
    volume = np.ascontiguousarray(volume.transpose(2, 1, 0))
    resolutions = np.abs(hdr["pixdim"][1:4]).astype(np.float32)
    affine = hdr["affine"]
    if np.any(np.isnan(affine)) and hdr["qform"] is not None:
        affine = hdr["qform"]
    return volume, resolutions, affine


def save_slices(folder: str, images: List[Slice]) -> None:
    os.makedirs(folder, exist_ok=True)
    for i, image in enumerate(images):
        image.save(os.path.join(folder, f"{i}.nii.gz"), True)


def load_slices(folder: str, device=torch.device("cpu")) -> List[Slice]:
    """`<id>.nii[.gz]` files of a folder -> slices ordered by id; mask = image > 0 (image.py:305-330)."""
    slices, ids = [], []
    for f in os.listdir(folder):
        if not (f.endswith("nii") or f.endswith("nii.gz")):
            continue
        ids.append(int(f.split(".nii")[0]))
        image, resolutions, affine = load_nii_volume(os.path.join(folder, f))
        image_t = torch.tensor(image, device=device)
        image_t, mask_t, transformation = affine2transformation(image_t, image_t > 0, resolutions, affine)
        slices.append(Slice(image=image_t, mask=mask_t, transformation=transformation, resolution_x=float(resolutions[0]),
                            resolution_y=float(resolutions[1]), resolution_z=float(resolutions[2])))
    return [s for _, s in sorted(zip(ids, slices), key=lambda p: p[0])]


def _load_pair(path_vol: str, path_mask: Optional[str]):
    image, resolutions, affine = load_nii_volume(path_vol)
    if path_mask is None:
        mask = image > 0
    else:
        m, resolutions_m, affine_m = load_nii_volume(path_mask)
        mask = m > 0
        if not compare_resolution_affine(resolutions, affine, resolutions_m, affine_m, image.shape, mask.shape):
            raise Exception("Error: the sizes/resolutions/affine transformations of the input stack and stack mask do not match!")
    return image, mask, resolutions, affine


def load_stack(path_vol: str, path_mask: Optional[str] = None, device=torch.device("cpu")) -> Stack:
    """A NIfTI volume read as a stack of its z-slices, one transform per slice (image.py:333-365)."""
    image, mask, resolutions, affine = _load_pair(path_vol, path_mask)
    image_t, mask_t, transformation = affine2transformation(torch.tensor(image, device=device), torch.tensor(mask, device=device), resolutions, affine)
    return Stack(slices=image_t.unsqueeze(1), mask=mask_t.unsqueeze(1), transformation=transformation, resolution_x=float(resolutions[0]),
                 resolution_y=float(resolutions[1]), thickness=float(resolutions[2]), gap=float(resolutions[2]))


def load_volume(path_vol: str, path_mask: Optional[str] = None, device=torch.device("cpu")) -> Volume:
    """A NIfTI volume with ONE volume-centred transform (image.py:368-400).  The reference averages the per-slice
    axis-angle rows; they share one rotation and differ only in t_z, so the mean is that rotation with the mean translation
    -- taken here directly on the matrices (no converter round trip, works for CPU tensors too)."""
    image, mask, resolutions, affine = _load_pair(path_vol, path_mask)
    image_t, mask_t, transformation = affine2transformation(torch.tensor(image, device=device), torch.tensor(mask, device=device), resolutions, affine)
    mats = transformation.matrix(trans_first=True)
    centre = torch.cat([mats[:1, :, :3], mats[:, :, 3:].mean(0, keepdim=True)], -1)
    return Volume(image=image_t, mask=mask_t, transformation=RigidTransform(centre, trans_first=True), resolution_x=float(resolutions[0]),
                  resolution_y=float(resolutions[1]), resolution_z=float(resolutions[2]))
