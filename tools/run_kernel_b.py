#!/usr/bin/env python
"""Runs the slice-acquisition family (kernel B) once per operator on the BASELINE config-2 stacks -- the target of
`ncu --set full -k regex:"forward_kernel|backward_kernel"` captures (profiles/) -- and prints CUDA-event timings per operator.

    python tools/run_kernel_b.py [--reps 5]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch

    from nesvor_b200.csrc import build as nsv_build

    nsv_build.build()
    from nesvor_b200.data.phantom import STACK_ORIENTATIONS, phantom3d, stack_axisangles, stack_geometry
    from nesvor_b200.slice_acquisition import slice_acquisition, slice_acquisition_adjoint
    from nesvor_b200.transform import RigidTransform, mat_update_resolution
    from nesvor_b200.utils import get_PSF

    dev = torch.device("cuda", 0)
    n = 128
    ss, n_slice = stack_geometry(n, 1.0, 1.0, 3.0)
    vol = torch.tensor(phantom3d(n), dtype=torch.float32, device=dev)[None, None]
    psf = get_PSF(res_ratio=(1.0, 1.0, 3.0), device=dev)
    ax = stack_axisangles(STACK_ORIENTATIONS[:3], n_slice, 3.0).to(dev)
    mat = mat_update_resolution(RigidTransform(ax, trans_first=True).matrix(), 1, 1.0).contiguous()
    taps = int((psf != 0).sum())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn):
        durs = []
        for i in range(1 + a.reps):
            flush.zero_()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            out = fn()
            k1.record()
            torch.cuda.synchronize()
            if i >= 1:
                durs.append(k0.elapsed_time(k1))
        return out, (sum(durs) / len(durs) if durs else float('nan'))

    slices, t_fwd = timed(lambda: slice_acquisition(mat, vol, None, None, psf, (ss, ss), 1.0, False, False))
    n_px = slices.numel()
    _, t_adj = timed(lambda: slice_acquisition_adjoint(mat, psf, slices, None, None, vol.shape[-3:], 1.0, False, False))
    _, t_adj_eq = timed(lambda: slice_acquisition_adjoint(mat, psf, slices, None, None, vol.shape[-3:], 1.0, False, True))

    def fwd_bwd():
        v = vol.clone().requires_grad_(True)
        m = mat.clone().requires_grad_(True)
        s = slice_acquisition(m, v, None, None, psf, (ss, ss), 1.0, False, False)
        s.backward(slices)
        return v.grad

    _, t_fb = timed(fwd_bwd)
    alg = n_px * (taps * 8 * 4 + 4)
    print(json.dumps({"slices": int(slices.shape[0]), "slice_shape": [ss, ss], "psf_taps": taps, "pixels": n_px,
                      "algorithmic_bytes_per_pass": alg, "ms": {"forward": t_fwd, "adjoint": t_adj, "adjoint_equalized": t_adj_eq,
                                                                 "forward+backward(vol,transforms)": t_fb},
                      "forward_GBps": alg / (t_fwd * 1e-3) / 1e9, "adjoint_GBps": alg / (t_adj * 1e-3) / 1e9}))


if __name__ == "__main__":
    main()
