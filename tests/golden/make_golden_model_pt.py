#!/usr/bin/env python
"""Writes tests/golden/reference_model.pt with the REFERENCE'S OWN writer -- `nesvor.cli.io.outputs` (cli/io.py:33-49), the
reference's own `INR`, `Volume` and `RigidTransform` classes (so the pickle carries the reference's module paths) -- running
on nesvor_b200.compat (tiny-cuda-nn is absent: the two tcnn modules are this package's stand-ins, i.e. the flat parameter
layout is ours), and tests/golden/reference_model_tensors.npz with the tensors that went in.
`nesvor_b200.io.load_model` must read the file back (tests/test_io_host.py).

    python tests/golden/make_golden_model_pt.py [--out DIR]     # needs /root/reference or baseline/_ref
"""
import argparse
import os
import sys
from argparse import Namespace

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=HERE)
    a = ap.parse_args()
    import numpy as np
    import torch

    sys.path.insert(0, ROOT)
    ref = next((p for p in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")) if os.path.isdir(os.path.join(p, "nesvor"))), None)
    if ref is None:
        raise SystemExit("no copy of the reference package found")
    sys.path.insert(0, ref)
    import nesvor_b200.compat as compat

    compat.install(fused=False)
    import nesvor.cli.io as rio
    import nesvor.nesvor.models as rm
    from nesvor.image import Volume
    from nesvor.transform import RigidTransform

    args = Namespace(n_features_per_level=2, log2_hashmap_size=12, level_scale=2.0, coarsest_resolution=16.0, finest_resolution=4.0, depth=1,
                     width=64, n_features_z=15, single_precision=False, dtype=torch.float16, device=torch.device("cpu"), output_volume=None,
                     output_slices=None, simulated_slices=None, output_intensity_mean=None, output_model=os.path.join(a.out, "reference_model.pt"))
    torch.manual_seed(20261017)
    inr = rm.INR(torch.tensor([[-30.0, -28.0, -26.0], [30.0, 31.0, 32.0]]), args)
    with torch.no_grad():
        inr.encoding.params.uniform_(-1, 1)
        inr.density_net.params.uniform_(-0.5, 0.5)
    g = torch.Generator().manual_seed(1)
    img = torch.rand(4, 5, 6, generator=g)
    mask = Volume(img, img > 0.3, RigidTransform(torch.tensor([[0.1, -0.2, 0.3, 1.0, 2.0, 3.0]])), 0.8, 0.8, 0.8)
    rio.outputs({"output_model": inr, "mask": mask}, args)
    np.savez(os.path.join(a.out, "reference_model_tensors.npz"), mask_image=img.numpy(), mask_mask=(img > 0.3).numpy(),
             mask_axisangle=mask.transformation.axisangle().numpy(), **{k: v.float().numpy() for k, v in inr.state_dict().items()})
    print("wrote", args.output_model, os.path.getsize(args.output_model), "bytes; writer:", rio.__file__)


if __name__ == "__main__":
    main()
