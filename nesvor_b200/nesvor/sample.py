"""Inference: PSF-averaged sampling of the INR into volumes and slices.

Host-side mirror of nesvor/nesvor/sample.py:10-64 (`sample_volume`, `sample_points`, `sample_slice`,
`sample_slices`).  When `args.fused` is set the batched render goes through the forward-only fused
kernel `nsv_inr_render` instead of sample_batch + INR.forward.
"""
from argparse import Namespace
from typing import List

import torch

from ..image import Slice, Volume
from ..transform import transform_points
from ..utils import meshgrid, resolution2sigma
from .models import INR


def _render(model: INR, xyz, transformation, psf_sigma, n_samples: int, args: Namespace) -> torch.Tensor:
    if getattr(args, "fused", False):
        from .fused import attach_render_state, fused_render

        if getattr(model, "_fused_state", None) is None:
            attach_render_state(model, args)  # snapshot of the (trained) parameters in kernel layout
        return fused_render(model, xyz, transformation, psf_sigma, n_samples)
    xyz_batch = model.sample_batch(xyz, transformation, psf_sigma, n_samples)
    return model(xyz_batch, False).mean(-1)


def sample_volume(model: INR, mask: Volume, args: Namespace) -> Volume:
    model.eval()
    img = mask.resample(args.output_resolution, None)
    img.image[img.mask] = sample_points(model, img.xyz_masked, args)
    return img


def sample_points(model: INR, xyz: torch.Tensor, args: Namespace) -> torch.Tensor:
    shape = xyz.shape[:-1]
    xyz = xyz.view(-1, 3)
    v = torch.empty(xyz.shape[0], dtype=torch.float32, device=args.device)
    batch_size = args.inference_batch_size
    n = 0 if args.no_output_psf else args.n_inference_samples
    with torch.no_grad():
        for i in range(0, xyz.shape[0], batch_size):
            v[i : i + batch_size] = _render(model, xyz[i : i + batch_size], None,
                                            resolution2sigma(args.output_resolution, isotropic=True), n, args)
    return v.view(shape)


def sample_slice(model: INR, slice: Slice, mask: Volume, args: Namespace) -> Slice:
    slice_sampled = slice.clone()
    slice_sampled.image = torch.zeros_like(slice_sampled.image)
    slice_sampled.mask = torch.zeros_like(slice_sampled.mask)
    xyz = meshgrid(slice_sampled.shape_xyz, slice_sampled.resolution_xyz).view(-1, 3)
    m = mask.sample_points(transform_points(slice_sampled.transformation, xyz)) > 0
    if m.any():
        n = 0 if args.no_output_psf else args.n_inference_samples
        v = _render(model, xyz[m], slice_sampled.transformation, resolution2sigma(slice_sampled.resolution_xyz, isotropic=False), n, args)
        slice_sampled.mask = m.view(slice_sampled.mask.shape)
        slice_sampled.image[slice_sampled.mask] = v.to(slice_sampled.image.dtype)
    return slice_sampled


def sample_slices(model: INR, slices: List[Slice], mask: Volume, args: Namespace) -> List[Slice]:
    model.eval()
    with torch.no_grad():
        return [sample_slice(model, s, mask, args) for s in slices]
