from .image import (Image, Slice, Stack, Volume, load_nii_volume, load_slices, load_stack, load_volume, save_nii_volume,
                    save_slices)
from .affine import affine2transformation, compare_resolution_affine, transformation2affine
