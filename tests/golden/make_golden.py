"""Generates the committed golden fixtures from the REFERENCE itself (run in the build container,
where /root/reference exists):

  python tests/golden/make_golden.py

* slice_acq_ref.npz / pose_ref.npz  outputs of the reference's own kernel bodies
  (slice_acq_cuda_kernel.cu:18-950, transform_convert_cuda_kernel.cu:15-440) executed on CPU
  through oracle/_ref (oracle/build_ref.sh), single thread, for the seeded inputs of
  tests/helpers.py:slice_acq_case / the reference's 11 axis-angle vectors.
* psf_ref.npz      nesvor.utils.psf.get_PSF / resolution2sigma imported from /root/reference.
* phantom_ref.npz  tests/phantom3d.py imported from /root/reference (n = 16, 32) + sha1 of n = 64.
The fixtures are small (< 1 MB in total) and let the GPU box, which has no /root/reference, check
the oracle and the CUDA kernels against reference outputs.
"""
import hashlib
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("NSV_REFERENCE_ROOT", "/root/reference")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    from oracle import native
    from helpers import REF_AXISANGLES, slice_acq_case

    native.set_threads(1)
    ref = native.Reference()
    assert ref is not None, "needs /root/reference"

    out = {}
    for tag, kw in (("plain", dict(masks=False)), ("masked", dict(masks=True, seed=1))):
        for interp in (0, 1):
            c = slice_acq_case(**kw)
            key = f"{tag}_i{interp}"
            fwd = ref.forward(c["transforms"], c["vol"], c["vol_mask"], c["slices_mask"], c["psf"], c["slice_shape"], c["res_slice"], True, interp)
            out[key + "_slices"], out[key + "_weight"] = fwd
            gv, gt = ref.backward(c["transforms"], c["vol"], c["vol_mask"], c["psf"], c["grad_slices"], c["slices_mask"], c["res_slice"], interp, True, True)
            out[key + "_bwd_grad_vol"], out[key + "_bwd_grad_tf"] = gv, gt
            for eq in (0, 1):
                vol, vw = ref.adjoint_forward(c["transforms"], c["psf"], c["slices"], c["slices_mask"], c["vol_mask"], c["vol_shape"], c["res_slice"], interp, eq)
                out[f"{key}_adj{eq}_vol"] = vol
                gs, gt2 = ref.adjoint_backward(c["transforms"], c["grad_vol"].copy(), vw, c["vol_mask"], c["psf"], c["slices"], c["slices_mask"], vol, c["res_slice"], interp, eq, True, True)
                out[f"{key}_adjbwd{eq}_grad_slices"], out[f"{key}_adjbwd{eq}_grad_tf"] = gs, gt2
    np.savez_compressed(os.path.join(HERE, "slice_acq_ref.npz"), **{k: v.astype(np.float32) for k, v in out.items()})

    ax = np.array(REF_AXISANGLES, np.float32)
    rng = np.random.default_rng(7)
    extra = rng.normal(size=(32, 6)).astype(np.float32)
    extra[:4, :3] *= 1e-4
    ax_all = np.concatenate([ax, extra])
    mat = ref.axisangle2mat_forward(ax_all)[0]
    g = rng.normal(size=mat.shape).astype(np.float32)
    ga = rng.normal(size=ax_all.shape).astype(np.float32)
    np.savez_compressed(
        os.path.join(HERE, "pose_ref.npz"), axisangle=ax_all, mat=mat, grad_mat=g, grad_axisangle=ga,
        a2m_bwd=ref.axisangle2mat_backward(g, ax_all)[0], m2a_fwd=ref.mat2axisangle_forward(mat)[0],
        m2a_bwd=ref.mat2axisangle_backward(mat, ga)[0])

    psf_mod = _load(os.path.join(REF, "nesvor/utils/psf.py"), "ref_psf")
    psfs = {}
    for ratio in ((1.5, 1.5, 3.0), (1.25, 1.25, 3.75), (1.0, 1.0, 3.0), (1.0, 1.0, 1.0)):
        psfs["psf_%g_%g_%g" % ratio] = psf_mod.get_PSF(res_ratio=ratio).numpy()
    psfs["constants"] = np.array([psf_mod.GAUSSIAN_FWHM, psf_mod.SINC_FWHM])
    np.savez_compressed(os.path.join(HERE, "psf_ref.npz"), **psfs)

    ph = _load(os.path.join(REF, "tests/phantom3d.py"), "ref_phantom")
    np.savez_compressed(
        os.path.join(HERE, "phantom_ref.npz"), n16=ph.phantom3d(n=16).astype(np.float32), n32=ph.phantom3d(n=32).astype(np.float32),
        sha1_n64=np.frombuffer(hashlib.sha1(ph.phantom3d(n=64).astype(np.float32).tobytes()).digest(), np.uint8))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
