from .srr import CG, SRR, PSFreconstruction  # noqa: F401
