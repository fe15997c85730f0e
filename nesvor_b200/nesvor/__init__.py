from .models import INR, INRModel, NeSVoR, build_encoding, build_network, tv_reg, edge_reg, l2_reg
from .models import D_LOSS, S_LOSS, DS_LOSS, B_REG, T_REG, I_REG
from .train import train, Dataset
from .sample import sample_volume, sample_points, sample_slice, sample_slices
