"""The two consumers of kernel B in the reference's SVoRT driver that sit next to the INR path
(nesvor/svort/inference.py:370-444): the volume reconstructed from (registered) stacks that
initialises everything downstream, and the slice-wise similarity between acquired slices and slices simulated from a
volume.  Registration itself (the SVoRT transformer, VVR) is out of scope.
"""
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F

from ..slice_acquisition import slice_acquisition
from ..transform import RigidTransform, mat_update_resolution
from ..utils import get_PSF
from ..utils.loss import ncc_loss
from .srr import SRR, PSFreconstruction

VOLUME_SHAPE = (256, 256, 256)  # inference.py:395: SVoRT's fixed reconstruction grid


def _pad_to_square(stacks: List[torch.Tensor]) -> List[torch.Tensor]:
    """Zero-pads every stack [n, 1, h, w] symmetrically (extra pixel at the end) to the largest in-plane size of all."""
    size = max(max(s.shape[-2:]) for s in stacks)
    out = []
    for s in stacks:
        dy, dx = size - s.shape[-2], size - s.shape[-1]
        out.append(F.pad(s, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2)) if dx > 0 or dy > 0 else s)
    return out


def reconstruct_from_stacks(transforms: List[RigidTransform], stacks: List[torch.Tensor], res_s: float, s_thick: float, res_r: float,
                            n_stack_recon: Optional[int], volume_shape=VOLUME_SHAPE) -> torch.Tensor:
    """PSF-weighted scatter of the first `n_stack_recon` stacks (all when None) followed by ONE conjugate-gradient SRR
    iteration on the pixels that carry signal (inference.py:370-409).  `transforms[j]` are in millimetres; kernel B
    wants voxel units of the reconstruction grid (`mat_update_resolution`)."""
    padded = _pad_to_square(stacks)
    n = len(padded) if n_stack_recon is None else n_stack_recon
    params = {"psf": get_PSF(res_ratio=(res_s / res_r, res_s / res_r, s_thick / res_r), device=stacks[0].device),
              "slice_shape": padded[0].shape[-2:], "interp_psf": False, "res_s": res_s, "res_r": res_r, "s_thick": s_thick,
              "volume_shape": tuple(volume_shape)}
    mat = mat_update_resolution(RigidTransform.cat([transforms[j] for j in range(n)]).matrix(), 1, res_r)
    slices = torch.cat(padded[:n])
    volume = PSFreconstruction(mat, slices, None, None, params)
    return SRR(n_iter=1, use_CG=True)(mat, slices, volume, params, slices_mask=slices > 0)


def simulated_ncc(transforms: List[RigidTransform], stacks: List[torch.Tensor], volume: torch.Tensor, res_s: float, s_thick: float,
                  res_r: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per slice: global NCC between the acquired slice and the slice simulated from `volume` through kernel B inside the
    slice's own mask, and the mask sizes as weights (inference.py:412-444)."""
    psf = get_PSF(res_ratio=(res_s / res_r, res_s / res_r, s_thick / res_r), device=stacks[0].device)
    ncc, weight = [], []
    for stack, transform in zip(stacks, transforms):
        mask = stack > 0
        simulated = slice_acquisition(mat_update_resolution(transform.matrix(), 1, res_r), volume, None, mask, psf, stack.shape[-2:],
                                      res_s / res_r, False, False)
        weight.append(mask.sum((1, 2, 3)))
        ncc.append(ncc_loss(simulated, stack, mask, win=None, reduction="none"))
    ncc_all = torch.cat(ncc)
    return ncc_all, torch.cat(weight).view(ncc_all.shape)
