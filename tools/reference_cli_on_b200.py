#!/usr/bin/env python
"""`nesvor reconstruct` -- the reference's OWN command line, unmodified -- on libnesvor_b200 (GPU box).

baseline/_ref/nesvor is the reference's Python layer exactly as under /root/reference; nesvor_b200.compat supplies its
native imports (slice_acq_cuda, transform_convert_cuda, tinycudann) and, when nibabel is absent, the three nibabel calls it
makes.  This script

  1. simulates motion-corrupted stacks of the 3-D Shepp-Logan phantom with this package's simulator (kernel B) and writes
     them as a NIfTI slice folder (each slice with its true pose, as after registration);
  2. calls `nesvor.cli.main.main()` with
         nesvor reconstruct --input-slices <folder> --output-volume <out.nii.gz> --output-model <model.pt> --n-iter ... 
     i.e. the reference's argument parser, `inputs()`, `train()`, `sample_volume()`, `sample_slices()`, `outputs()`;
  3. reads the written volume back and scores it against the phantom at the volume's own voxel positions.

Prints ONE JSON line ({"available": false, ...} without the reference copy or a GPU).
"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
REF_PARENT = os.path.join(ROOT, "baseline", "_ref")


def main():
    import torch

    if not os.path.isdir(os.path.join(REF_PARENT, "nesvor")) or not torch.cuda.is_available():
        print(json.dumps({"available": False, "why": "baseline/_ref/nesvor absent" if torch.cuda.is_available() else "no CUDA device"}))
        return
    import nesvor_b200 as nb
    import nesvor_b200.compat as compat
    import psnr_phantom as pp
    from nesvor_b200.data.phantom import simulate_slices, stack_geometry

    dev = torch.device("cuda", 0)
    n, n_stacks, gap = 48, 3, 2.0
    torch.manual_seed(0)
    slices, volume, true_ax = simulate_slices(device=dev, n=n, n_stacks=n_stacks, res_r=1.0, res_s=1.0, gap=gap, motion_deg=3.0, motion_mm=1.5)
    _, n_slice = stack_geometry(n, 1.0, 1.0, gap)
    for s in slices:  # hand over the poses the data were acquired at (what registration would have produced)
        i = s.stack_idx * n_slice + s.slice_idx
        s.transformation = nb.RigidTransform(true_ax[i : i + 1].clone(), True)
    tmp = tempfile.mkdtemp(prefix="nsv_cli_")
    folder, out_vol, out_model = os.path.join(tmp, "slices"), os.path.join(tmp, "volume.nii.gz"), os.path.join(tmp, "model.pt")
    nb.save_slices(folder, slices)

    sys.path.insert(0, REF_PARENT)
    compat.install()  # native stand-ins + `train` rebound to the fused iteration (compat.fused_train)
    import nesvor.cli.main as cli  # the reference's command line
    import nesvor.nesvor.train as rtrain

    n_iter = 1500
    argv = ["nesvor", "reconstruct", "--input-slices", folder, "--output-volume", out_vol, "--output-model", out_model, "--n-iter", str(n_iter),
            "--batch-size", "2048", "--n-samples", "64", "--output-resolution", "1.0", "--verbose", "0", "--seed", "0"]
    old_argv = sys.argv
    sys.argv = argv
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    try:
        cli.main()
    finally:
        sys.argv = old_argv
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    train_info = dict(compat.LAST_TRAIN_INFO)
    # ---- score the written volume against the phantom at its own voxel positions
    rec = nb.load_volume(out_vol, device=dev)
    phantom = nb.Volume(volume[0, 0], volume[0, 0] > -1, nb.RigidTransform(torch.zeros(1, 6, device=dev), True), 1.0, 1.0, 1.0)
    m = rec.mask
    xyz = rec.xyz_masked
    gt = phantom.sample_points(xyz)
    inside = gt > 0
    out = {"available": True, "argv": argv[1:], "reference_train_file": rtrain.__file__, "n_slices": len(slices), "wall_s": wall,
           "queries": n_iter * 2048 * 64, "queries_per_s_wall_whole_command": n_iter * 2048 * 64 / wall,
           "volume_shape": list(rec.image.shape), "masked_voxels": int(m.sum()), "finite": bool(torch.isfinite(rec.image).all()),
           "psnr_inside": pp.psnr(rec.image[m][inside].cpu(), gt[inside].cpu()), "model_written": os.path.exists(out_model),
           "train": train_info, "train_is_fused_adapter": bool(getattr(rtrain.train, "__nesvor_b200_fused__", False))}
    # ---- the same command line on the BASELINE config-2 workload (bench.py's): ms per iteration of the hot loop as the CLI runs it
    try:
        out["config2_through_cli"] = config2_through_cli(cli, compat, nb, dev, tmp)
    except Exception as e:  # the accuracy run above stands on its own
        out["config2_through_cli"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    print(json.dumps(out))


def config2_through_cli(cli, compat, nb, dev, tmp, n_iter=600):
    """BASELINE config 2 (phantom 128^3, 3 orthogonal stacks, L = 16, T = 2^19, 64 x 3 hidden, 8192 px x 128 samples) through
    `nesvor reconstruct`: what bench.py times through FusedTrainer.step, timed here inside the command the user runs."""
    import torch

    from nesvor_b200.data.phantom import simulate_slices

    torch.manual_seed(0)
    slices, _, _ = simulate_slices(device=dev, n=128, n_stacks=3, res_r=1.0, res_s=1.0, gap=3.0)
    folder = os.path.join(tmp, "slices_cfg2")
    nb.save_slices(folder, slices)
    # the reference derives the number of levels from the data's bounding box (models.py:79-101):
    # n_levels = ceil(log2(extent / finest / base) / log2(scale) + 1), base = ceil(extent / coarsest).  BASELINE config 2
    # names L = 16, so --finest-resolution is chosen to land in the middle of the L = 16 bracket for THIS folder's extent.
    import math
    from argparse import Namespace

    from nesvor_b200.nesvor.train import Dataset

    bb = Dataset(nb.load_slices(folder, dev), Namespace(mask_threshold=1.0)).bounding_box
    extent = float((bb[1] - bb[0]).max())
    base = math.ceil(extent / 16.0)
    finest = extent / base / 1.3819**14.5
    argv = ["nesvor", "reconstruct", "--input-slices", folder, "--output-model", os.path.join(tmp, "model_cfg2.pt"), "--n-iter", str(n_iter),
            "--batch-size", "8192", "--n-samples", "128", "--depth", "3", "--finest-resolution", "%.6f" % finest, "--no-pixel-variance",
            "--no-slice-variance", "--no-transformation-optimization", "--verbose", "0", "--seed", "0"]
    # wall time of the command's phases: the names `nesvor.cli.commands` calls (cli/commands.py:100-125) wrapped with timers
    import time as _time

    import nesvor.cli.commands as cmds

    phases = {}

    def timed(name, fn):
        def call(*a_, **k_):
            torch.cuda.synchronize()
            t0 = _time.perf_counter()
            try:
                return fn(*a_, **k_)
            finally:
                torch.cuda.synchronize()
                phases[name] = phases.get(name, 0.0) + _time.perf_counter() - t0

        return call

    saved = {n: getattr(cmds, n) for n in ("inputs", "train", "sample_volume", "sample_slices", "outputs")}
    for n, fn in saved.items():
        setattr(cmds, n, timed(n, fn))
    old_argv, sys.argv = sys.argv, argv
    torch.cuda.synchronize()
    t_cmd = _time.perf_counter()
    try:
        cli.main()
    finally:
        sys.argv = old_argv
        for n, fn in saved.items():
            setattr(cmds, n, fn)
    torch.cuda.synchronize()
    info = dict(compat.LAST_TRAIN_INFO)
    info["command_wall_s"] = _time.perf_counter() - t_cmd
    info["phase_wall_s"] = phases
    info["argv"] = argv[1:]
    if "ms_per_iteration" in info:
        info["queries_per_s"] = info["queries_per_iteration"] / info["ms_per_iteration"] * 1e3
    return info


if __name__ == "__main__":
    main()
