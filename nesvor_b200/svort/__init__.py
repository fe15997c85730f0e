from .srr import CG, SRR, PSFreconstruction  # noqa: F401
from .inference import reconstruct_from_stacks, simulated_ncc  # noqa: F401
