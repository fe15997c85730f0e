"""Golden vectors for the NIfTI-affine geometry, produced by the REFERENCE's own functions
(nesvor/image/image_utils.py: affine2transformation, transformation2affine, compare_resolution_affine), run in the build
container where /root/reference exists:

  python tests/golden/make_golden_affine.py        ->  tests/golden/affine_ref.npz, tests/golden/ncc_ref.npz

image_utils.py imports nibabel (absent here) and `..transform` (whose import JIT-compiles a CUDA extension), but the three
functions only need numpy, torch and a container with `.matrix(trans_first=True)`: both imports are satisfied with stubs
and the reference file is executed unmodified from where it lies.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NSV_REFERENCE_ROOT", "/root/reference")


class _Rigid:  # what image_utils.py touches of nesvor.transform.RigidTransform
    def __init__(self, data, trans_first=True):
        assert trans_first and data.ndim == 3
        self.mat = data

    def matrix(self, trans_first=True):
        assert trans_first
        return self.mat


def load_reference_image_utils():
    sys.modules.setdefault("nibabel", types.ModuleType("nibabel"))
    pkg = types.ModuleType("refpkg")
    pkg.__path__ = []
    sub = types.ModuleType("refpkg.image")
    sub.__path__ = []
    tr = types.ModuleType("refpkg.transform")
    tr.RigidTransform = _Rigid
    sys.modules.update({"refpkg": pkg, "refpkg.image": sub, "refpkg.transform": tr})
    spec = importlib.util.spec_from_file_location("refpkg.image.image_utils", os.path.join(REF, "nesvor/image/image_utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def random_rotation(rng):
    q, r = np.linalg.qr(rng.normal(size=(3, 3)))
    q = q * np.sign(np.diag(r))
    if np.linalg.det(q) < 0:
        q[:, 2] = -q[:, 2]
    return q


def main():
    ref = load_reference_image_utils()
    rng = np.random.default_rng(11)
    out = {}
    n_cases = 8
    for i in range(n_cases):
        d, h, w = (int(v) for v in rng.integers(3, 9, size=3))
        res = rng.uniform(0.5, 4.0, size=3)
        M = random_rotation(rng) @ np.diag(res)
        if i % 3 == 2:  # left-handed voxel axes
            M[:, 0] = -M[:, 0]
        A = np.eye(4)
        A[:3, :3] = M
        A[:3, 3] = rng.uniform(-80, 80, size=3)
        vol = torch.tensor(rng.normal(size=(d, h, w)), dtype=torch.float32)
        mask = vol > 0
        v2, m2, tr = ref.affine2transformation(vol, mask, res.astype(np.float32), A)
        out[f"a2t_{i}_vol"], out[f"a2t_{i}_res"], out[f"a2t_{i}_affine"] = vol.numpy(), res.astype(np.float32), A
        out[f"a2t_{i}_vol_out"], out[f"a2t_{i}_mask_out"], out[f"a2t_{i}_mat"] = v2.numpy(), m2.numpy(), tr.matrix().numpy()
        # transformation2affine on an independent rigid transform of the same image
        T = torch.tensor(np.concatenate([random_rotation(rng), rng.uniform(-50, 50, size=(3, 1))], -1)[None], dtype=torch.float32)
        aff = ref.transformation2affine(vol, _Rigid(T.clone()), float(res[0]), float(res[1]), float(res[2]))
        out[f"t2a_{i}_mat"], out[f"t2a_{i}_affine"] = T.numpy(), aff
    out["n_cases"] = np.array(n_cases)
    r, a = np.array([1.0, 1.0, 3.0]), np.eye(4)
    out["compare"] = np.array([ref.compare_resolution_affine(r, a, r + 5e-4, a, (3, 4, 5), (3, 4, 5)),
                               ref.compare_resolution_affine(r, a, r + 2e-3, a, (3, 4, 5), (3, 4, 5)),
                               ref.compare_resolution_affine(r, a, r, a + 2e-3, (3, 4, 5), (3, 4, 5)),
                               ref.compare_resolution_affine(r, a, r, a, (3, 4, 5), (3, 4, 6))])
    np.savez_compressed(os.path.join(HERE, "affine_ref.npz"), **out)
    print("wrote affine_ref.npz:", len(out), "arrays")


def make_ncc():
    """ncc_ref.npz: nesvor/utils/loss.py::ncc_loss (pure torch, imported from the reference file as it lies)."""
    spec = importlib.util.spec_from_file_location("ref_loss", os.path.join(REF, "nesvor/utils/loss.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(21)
    I = torch.rand(5, 1, 12, 14, generator=g)
    J = 0.6 * I + 0.4 * torch.rand(5, 1, 12, 14, generator=g)
    mask = torch.rand(5, 1, 12, 14, generator=g) > 0.3
    V, W = torch.rand(2, 2, 6, 7, 8, generator=g), torch.rand(2, 2, 6, 7, 8, generator=g)
    out = {"I": I.numpy(), "J": J.numpy(), "mask": mask.numpy(), "V": V.numpy(), "W": W.numpy(),
           "global_masked": mod.ncc_loss(I, J, mask, win=None, reduction="none").numpy(),
           "global_plain": mod.ncc_loss(I, J, None, win=None, reduction="none").numpy(),
           "win9": mod.ncc_loss(I, J, None, win=9).numpy(), "win9_level1_masked": mod.ncc_loss(I, J, mask, win=9, level=1).numpy(),
           "win5_3d_mean": mod.ncc_loss(V, W, None, win=5, reduction="mean").numpy(),
           "win5_3d_sum": mod.ncc_loss(V, W, None, win=5, reduction="sum").numpy()}
    np.savez_compressed(os.path.join(HERE, "ncc_ref.npz"), **out)
    print("wrote ncc_ref.npz")


if __name__ == "__main__":
    main()
    make_ncc()
