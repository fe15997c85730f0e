"""The zero-edit drop-in (nesvor_b200/compat.py): with the three stand-in modules installed, the UNMODIFIED reference
package imports and builds its own INR / NeSVoR on this library's modules, and its native calls land in the C ABI.
Runs where /root/reference exists (the build container); a subprocess keeps the reference package out of this process.
No compute: there is no GPU here -- the native entry points are reached and refuse CPU tensors like CHECK_CUDA does."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NSV_REFERENCE_ROOT", "/root/reference")

SCRIPT = r'''
import sys, types, json
from argparse import Namespace
import torch
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
import nesvor_b200.compat as compat
installed = compat.install()                         # nibabel is absent in this image: its stand-in is installed too
import nesvor                                        # the reference package, as it lies under /root/reference
import nesvor.slice_acquisition.slice_acq as rsa
import nesvor.transform.transform_convert as rtc
import nesvor.nesvor.models as rm
import nesvor.nesvor.train as rtrain
from nesvor.transform import RigidTransform
import nesvor_b200 as nb
out = {"installed": sorted(installed)}
out["sa_is_ours"] = rsa.slice_acq_cuda is sys.modules["nesvor.slice_acq_cuda"] and rsa.slice_acq_cuda.forward.__module__.startswith("nesvor_b200")
out["tc_is_ours"] = rtc.transform_convert_cuda.axisangle2mat_forward.__module__.startswith("nesvor_b200")
out["ref_file"] = rm.__file__
args = Namespace(n_features_per_level=2, log2_hashmap_size=19, level_scale=1.3819, coarsest_resolution=16.0, finest_resolution=0.5,
                 n_levels_bias=4, depth=1, width=64, n_features_z=15, n_features_slice=16, no_transformation_optimization=False,
                 no_slice_scale=False, no_pixel_variance=False, no_slice_variance=False, single_precision=False, dtype=torch.float16,
                 image_regularization="edge", delta=0.2, device=torch.device("cpu"))
bb = torch.tensor([[-55.0, -55.0, -55.0], [55.0, 55.0, 55.0]])
ref_inr = rm.INR(bb, args)                           # the reference's class, built on tcnn.Encoding / tcnn.Network = ours
our_inr = nb.INR(bb, args)
out["encoding_class"] = type(ref_inr.encoding).__mro__[1].__name__
out["network_class"] = type(ref_inr.density_net).__mro__[1].__name__
out["state_keys"] = sorted(ref_inr.state_dict())
out["same_shapes"] = {k: list(v.shape) == list(our_inr.state_dict()[k].shape) for k, v in ref_inr.state_dict().items()}
out["n_levels"] = ref_inr.encoding.n_levels
ax = torch.randn(6, 6) * 0.1
model = rm.NeSVoR(RigidTransform(ax), torch.tensor([[1.0, 1.0, 3.0]]).repeat(6, 1), 0.7, bb, args)
out["nesvor_modules"] = sorted(n for n, _ in model.named_children())
out["b_net_params"] = int(model.b_net.params.numel())
# native calls are reached and refuse CPU tensors (the reference's own wrappers, our entry points)
errs = {}
for name, call in (("axisangle2mat", lambda: rtc.axisangle2mat(ax)),
                   ("slice_acquisition", lambda: rsa.slice_acquisition(torch.zeros(2, 3, 4), torch.zeros(1, 1, 4, 4, 4), None, None, torch.ones(3, 3, 3), (5, 5), 1.0, False, False)),
                   ("encoding", lambda: ref_inr.encoding(torch.rand(8, 3)))):
    try:
        call(); errs[name] = "ran"
    except RuntimeError as e:
        errs[name] = str(e)[:60]
out["errors"] = errs
out["train_uses_ref_models"] = rtrain.NeSVoR is rm.NeSVoR
# the hot loop: `train` is rebound to the fused adapter where the reference looks it up, the original is kept for the fallback
import nesvor.cli.commands as cmds
out["train_rebound"] = bool(getattr(rtrain.train, "__nesvor_b200_fused__", False)) and cmds.train is rtrain.train
out["ref_train_kept"] = compat._REF_TRAIN is not None and compat._REF_TRAIN.__module__ == "nesvor.nesvor.train" and compat._REF_TRAIN is not rtrain.train
compat.install()  # idempotent: a second call must not wrap the adapter around itself
out["train_rebound_once"] = compat._REF_TRAIN.__module__ == "nesvor.nesvor.train" and cmds.train is rtrain.train
# ---- the reference's own image module on the nibabel stand-in: write with the reference, read with both
import os, tempfile
import numpy as np
import nesvor.image as rim
import nesvor_b200.image as oim
import nesvor.cli.main  # noqa: F401  the whole CLI imports
rng = np.random.default_rng(0)
q, r = np.linalg.qr(rng.normal(size=(3, 3)))
q = q * np.sign(np.diag(r))
if np.linalg.det(q) < 0:
    q[:, 2] = -q[:, 2]
mat = torch.tensor(np.concatenate([q, rng.uniform(-20, 20, (3, 1))], -1)[None], dtype=torch.float32)
img = torch.tensor(rng.uniform(0.1, 1, (4, 5, 6)), dtype=torch.float32)
tmp = tempfile.mkdtemp()
path = os.path.join(tmp, "v.nii.gz")
rim.Volume(img, img > 0.5, RigidTransform(mat), 0.9, 1.1, 3.0).save(path, masked=False)
a, b = rim.load_stack(path), oim.load_stack(path)
out["nifti"] = {"shim": bool(getattr(sys.modules["nibabel"], "__nesvor_b200_shim__", False)), "slices_equal": bool(torch.equal(a.slices, b.slices)),
                "image_survives": bool(torch.equal(a.slices[:, 0], img)),
                "transform_diff": float((a.transformation.matrix() - b.transformation.matrix(True)).abs().max()),
                "thickness": [float(a.thickness), float(b.thickness)]}
folder = os.path.join(tmp, "slices")
os.makedirs(folder)
rim.save_slices(folder, a[:])
back = oim.load_slices(folder)
out["nifti"]["slice_folder"] = len(back) == 4 and all(bool(torch.equal(x.image, y.image * y.mask)) for x, y in zip(back, a[:]))
print(json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "nesvor")), reason="needs the reference tree (build container only)")
def test_unmodified_reference_runs_on_the_standin_modules(native_lib):
    import json

    r = subprocess.run([sys.executable, "-c", f"ROOT={ROOT!r}; REF={REF!r}\n" + SCRIPT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert set(out["installed"]) >= {"nesvor.slice_acq_cuda", "nesvor.transform_convert_cuda", "tinycudann"}
    assert out["sa_is_ours"] and out["tc_is_ours"] and out["ref_file"].startswith(REF) and out["train_uses_ref_models"]
    assert "nesvor.nesvor.train.train" in out["installed"] and out["train_rebound"] and out["ref_train_kept"] and out["train_rebound_once"]
    assert out["encoding_class"] == "HashGridEncoding" and out["network_class"] == "FusedMLP"
    assert out["state_keys"] == ["bounding_box", "density_net.params", "encoding.params"] and all(out["same_shapes"].values())
    assert out["n_levels"] == 12  # 110 mm box, reference defaults (SURVEY s.8: base 7, L 12)
    assert {"inr", "sigma_net", "b_net", "slice_embedding"} <= set(out["nesvor_modules"]) and out["b_net_params"] == 64 * 32 + 16 * 64
    for k, v in out["errors"].items():
        assert "must be a CUDA tensor" in v, (k, v)
    nf = out["nifti"]  # files written by the reference's image.py (through the nibabel stand-in when nibabel is absent)
    assert nf["slices_equal"] and nf["image_survives"] and nf["transform_diff"] < 1e-5 and nf["thickness"] == [3.0, 3.0] and nf["slice_folder"]


def test_install_is_idempotent_and_reversible():
    import nesvor_b200.compat as compat

    before = {n: sys.modules.get(n) for n in ("nesvor.slice_acq_cuda", "nesvor.transform_convert_cuda", "tinycudann", "nibabel", "nibabel.nifti1")}
    try:
        a = compat.install()
        b = compat.install()
        assert set(a) == set(b) and hasattr(sys.modules["tinycudann"], "Encoding")
        assert compat.install(tcnn="never", nibabel="never").keys() == {"nesvor.slice_acq_cuda", "nesvor.transform_convert_cuda"}
        nib = compat.install(nibabel="force")["nibabel"]
        assert callable(nib.load) and callable(nib.save) and nib.nifti1.Nifti1Image is nib.Nifti1Image
        for fn in ("forward", "backward", "adjoint_forward", "adjoint_backward"):
            assert callable(getattr(sys.modules["nesvor.slice_acq_cuda"], fn))
        compat.uninstall()
        assert all(n not in sys.modules for n in before)
    finally:
        for n, m in before.items():
            if m is not None:
                sys.modules[n] = m
            else:
                sys.modules.pop(n, None)
