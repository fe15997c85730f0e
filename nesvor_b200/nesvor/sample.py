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

        st = getattr(model, "_fused_state", None)
        if st is None:
            st = attach_render_state(model, args)  # snapshot of the parameters in kernel layout
        elif getattr(st, "param_version", None) != model.encoding.params._version:
            st.pull_from_model()  # the INR was trained / loaded since the snapshot was taken
        st.param_version = model.encoding.params._version
        return fused_render(model, xyz, transformation, psf_sigma, n_samples)
    xyz_batch = model.sample_batch(xyz, transformation, psf_sigma, n_samples)
    return model(xyz_batch, False).mean(-1)


def _n_psf_samples(args: Namespace) -> int:
    return 0 if args.no_output_psf else args.n_inference_samples


def sample_volume(model: INR, mask: Volume, args: Namespace) -> Volume:
    """The INR rendered on the mask's grid resampled to `args.output_resolution` (sample.py:10-14)."""
    model.eval()
    out = mask.resample(args.output_resolution, None)
    out.image[out.mask] = sample_points(model, out.xyz_masked, args)
    return out


def sample_points(model: INR, xyz: torch.Tensor, args: Namespace) -> torch.Tensor:
    """PSF-averaged intensities at world points [..., 3], `args.inference_batch_size` points per launch, isotropic output
    PSF of `args.output_resolution` (sample.py:17-33)."""
    pts = xyz.reshape(-1, 3)
    sigma = resolution2sigma(args.output_resolution, isotropic=True)
    with torch.no_grad():
        vals = [_render(model, chunk, None, sigma, _n_psf_samples(args), args).to(torch.float32)
                for chunk in pts.split(args.inference_batch_size) if chunk.shape[0]]
    if not vals:
        return torch.empty(xyz.shape[:-1], dtype=torch.float32, device=args.device)
    return torch.cat(vals).view(xyz.shape[:-1])


def sample_slice(model: INR, slice: Slice, mask: Volume, args: Namespace) -> Slice:
    """The slice re-simulated from the INR at its own pose and resolution (anisotropic slice PSF), inside `mask` only;
    pixels outside the mask stay zero and unmasked (sample.py:36-53)."""
    out = slice.clone(zero=True)
    grid = meshgrid(out.shape_xyz, out.resolution_xyz).view(-1, 3)
    inside = mask.sample_points(transform_points(out.transformation, grid)) > 0
    if inside.any():
        sigma = resolution2sigma(out.resolution_xyz, isotropic=False)
        v = _render(model, grid[inside], out.transformation, sigma, _n_psf_samples(args), args)
        out.mask = inside.view(out.mask.shape)
        out.image[out.mask] = v.to(out.image.dtype)
    return out


def sample_slices(model: INR, slices: List[Slice], mask: Volume, args: Namespace) -> List[Slice]:
    model.eval()
    with torch.no_grad():
        return [sample_slice(model, s, mask, args) for s in slices]
