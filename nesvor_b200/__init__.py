"""nesvor_b200 -- B200-native (sm_100a) implementation of NeSVoR's reconstruction hot path.

Public surface = the reference's (daviddmc/NeSVoR @ f110505) names for this path:
    build_encoding, build_network, INR (= INRModel), NeSVoR, train, sample_volume, sample_points,
    sample_slice, sample_slices, slice_acquisition, slice_acquisition_adjoint, axisangle2mat,
    mat2axisangle, RigidTransform, get_PSF, resolution2sigma, (svort/srr.py) CG, SRR, PSFreconstruction, and (image/)
    Slice, Stack, Volume, load_stack, load_volume, load_slices, save_slices.
All compute runs in libnesvor_b200.so (nesvor_b200/csrc, C ABI in include/nesvor_b200.h); there is
no CPU or pure-PyTorch fallback.
"""
from .transform import (RigidTransform, axisangle2mat, mat2axisangle, mat_update_resolution, ax_update_resolution,
                        mat_transform_points, ax_transform_points, transform_points)
from .slice_acquisition import slice_acquisition, slice_acquisition_adjoint
from .utils import get_PSF, resolution2sigma
from .image import (Slice, Stack, Volume, load_slices, load_stack, load_volume, save_slices)
from .nesvor import (INR, INRModel, NeSVoR, build_encoding, build_network, train, Dataset, sample_volume, sample_points,
                     sample_slice, sample_slices)

from .svort import CG, SRR, PSFreconstruction

__version__ = "0.1.0"
