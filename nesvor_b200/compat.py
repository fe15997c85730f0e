"""Zero-edit drop-in: run the UNMODIFIED reference package (daviddmc/NeSVoR) on libnesvor_b200.

The reference reaches native code (and, for files, one third-party package) through these imports:

* `import nesvor.slice_acq_cuda`          (nesvor/slice_acquisition/slice_acq.py:5-19; else it JIT-compiles its .cu files)
* `import nesvor.transform_convert_cuda`  (nesvor/transform/transform_convert.py:3-18; same fallback)
* `import tinycudann as tcnn`             (nesvor/nesvor/models.py:7; `tcnn.Encoding` :25, `tcnn.Network` :31)

* `import nibabel as nib`                 (nesvor/image/image.py:5; `nib.load`, `nib.nifti1.Nifti1Image`, `nib.save` only)

`install()` registers modules under those names whose functions / classes have the pybind / tcnn signatures and
call the C ABI (include/nesvor_b200.h) -- `forward / backward / adjoint_forward / adjoint_backward`,
`axisangle2mat_{forward,backward} / mat2axisangle_{forward,backward}` (each returning a list of tensors, absent masks as
empty tensors, like slice_acq_cuda.cpp:61-161 and transform_convert_cuda.cpp:27-69), `Encoding(n_input_dims,
encoding_config, dtype)` and `Network(n_input_dims, n_output_dims, network_config)`.  After

    import nesvor_b200.compat as compat; compat.install()
    import nesvor                                   # the reference, untouched

the reference's own `INR`, `NeSVoR`, `train`, `slice_acquisition`, `RigidTransform`, SRR ... run on the B200 kernels
(the unfused path: one native op per reference op; the fused iteration is `nesvor_b200.train(..., args.fused=True)`).
A real `tinycudann` / `nibabel`, if installed, is left alone unless `tcnn="force"` / `nibabel="force"`; the nibabel stand-in
covers single-file NIfTI-1 (`image/nifti.py`).
"""
import importlib
import importlib.util
import sys
import types

import torch

_NAMES = ("nesvor.slice_acq_cuda", "nesvor.transform_convert_cuda", "tinycudann", "nibabel", "nibabel.nifti1")


def _lenient(fn):
    """The pybind modules insist on contiguous tensors (CHECK_CONTIGUOUS) and the reference's converter Functions hand
    autograd's gradients over as they come; newer torch versions produce expanded / strided gradients in places where the
    torch the reference was written for did not, so the stand-ins make tensor arguments contiguous instead of raising."""
    def call(*args):
        return fn(*[a.contiguous() if isinstance(a, torch.Tensor) and a.numel() and not a.is_contiguous() else a for a in args])

    call.__name__, call.__doc__, call.__module__ = fn.__name__, fn.__doc__, fn.__module__
    return call


def _slice_acq_module() -> types.ModuleType:
    sa = importlib.import_module("nesvor_b200.slice_acquisition.slice_acq")
    m = types.ModuleType("nesvor.slice_acq_cuda", "nesvor_b200 stand-in for the reference's slice_acq_cuda pybind module")
    for fn in ("forward", "backward", "adjoint_forward", "adjoint_backward"):
        setattr(m, fn, _lenient(getattr(sa, fn)))
    return m


def _transform_convert_module() -> types.ModuleType:
    tc = importlib.import_module("nesvor_b200.transform.transform_convert")
    m = types.ModuleType("nesvor.transform_convert_cuda", "nesvor_b200 stand-in for the reference's transform_convert_cuda pybind module")
    for fn in ("axisangle2mat_forward", "axisangle2mat_backward", "mat2axisangle_forward", "mat2axisangle_backward"):
        setattr(m, fn, _lenient(getattr(tc, fn)))
    return m


def _tcnn_module() -> types.ModuleType:
    from .nesvor.encoding import FusedMLP, HashGridEncoding

    class Encoding(HashGridEncoding):
        """tcnn.Encoding(n_input_dims, encoding_config, dtype=torch.float16): the HashGrid otype NeSVoR uses."""

        def __init__(self, n_input_dims, encoding_config, dtype=torch.float16, seed=1337):
            super().__init__(n_input_dims, dict(encoding_config), dtype, seed)

    class Network(FusedMLP):
        """tcnn.Network(n_input_dims, n_output_dims, network_config): ReLU MLP without biases, fp16 tensor cores."""

        def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
            super().__init__(n_input_dims, n_output_dims, dict(network_config), seed)

    m = types.ModuleType("tinycudann", "nesvor_b200 stand-in for the two tiny-cuda-nn modules NeSVoR instantiates")
    m.Encoding, m.Network = Encoding, Network
    m.__nesvor_b200_shim__ = True
    return m


def _nibabel_module() -> types.ModuleType:
    """The three nibabel calls the reference makes (nesvor/image/image.py:253-296) on image/nifti.py:
    `nib.load(path)` -> object with `.header["dim" | "pixdim"]`, `.get_fdata()`, `.affine`, `.get_qform()`;
    `nib.nifti1.Nifti1Image(data, affine)` with `.header.set_xyzt_units / set_qform / set_sform`; `nib.save(img, path)`."""
    import numpy as np

    from .image.nifti import read_nifti, write_nifti

    class _Header(dict):
        def set_xyzt_units(self, *a, **k):  # written files always carry millimetres
            pass

        def set_qform(self, affine, code=None):  # written files carry the affine as qform "aligned" ...
            self["_affine"] = np.asarray(affine, np.float64)

        def set_sform(self, affine, code=None):  # ... and as sform "scanner"
            self["_affine"] = np.asarray(affine, np.float64)

    class Nifti1Image:
        def __init__(self, dataobj, affine, header=None):
            self._data = np.asarray(dataobj)
            self.affine = np.eye(4) if affine is None else np.asarray(affine, np.float64)
            self.header = _Header(header or {})
            self._qform = None

        def get_fdata(self):
            return np.asarray(self._data, np.float64)

        def get_qform(self):
            return self.affine if self._qform is None else self._qform

    def load(path):
        data, hdr = read_nifti(str(path))
        img = Nifti1Image(data, hdr["affine"], {"dim": hdr["dim"], "pixdim": hdr["pixdim"].astype(np.float32)})
        img._qform = hdr["qform"]
        return img

    def save(img, path):
        write_nifti(str(path), img._data, img.header.get("_affine", img.affine))

    m = types.ModuleType("nibabel", "nesvor_b200 stand-in for the three nibabel calls NeSVoR makes (NIfTI-1 single files)")
    n1 = types.ModuleType("nibabel.nifti1")
    n1.Nifti1Image = Nifti1Image
    m.nifti1, m.Nifti1Image, m.load, m.save = n1, Nifti1Image, load, save
    m.__nesvor_b200_shim__ = n1.__nesvor_b200_shim__ = True
    return m


def _have_real(name: str) -> bool:
    cur = sys.modules.get(name)
    if cur is not None:
        return not getattr(cur, "__nesvor_b200_shim__", False)
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


def install(tcnn: str = "auto", nibabel: str = "auto") -> dict:
    """Registers the stand-in modules in `sys.modules`; returns {name: module} of what was installed.
    tcnn / nibabel = "auto": only when no real `tinycudann` / `nibabel` is importable; "force": always; "never": leave it alone."""
    done = {}
    for name, make in (("nesvor.slice_acq_cuda", _slice_acq_module), ("nesvor.transform_convert_cuda", _transform_convert_module)):
        sys.modules[name] = done[name] = make()
    if tcnn == "force" or (tcnn == "auto" and not _have_real("tinycudann")):
        sys.modules["tinycudann"] = done["tinycudann"] = _tcnn_module()
    if nibabel == "force" or (nibabel == "auto" and not _have_real("nibabel")):
        nib = _nibabel_module()
        sys.modules["nibabel"], sys.modules["nibabel.nifti1"] = nib, nib.nifti1
        done["nibabel"] = nib
    parent = sys.modules.get("nesvor")  # `import nesvor.x as y` resolves through the parent's attribute first
    if parent is not None:
        for name in ("slice_acq_cuda", "transform_convert_cuda"):
            setattr(parent, name, sys.modules["nesvor." + name])
    return done


def uninstall() -> None:
    for name in _NAMES:
        mod = sys.modules.get(name)
        if mod is not None and (not name.startswith(("tinycudann", "nibabel")) or getattr(mod, "__nesvor_b200_shim__", False)):
            del sys.modules[name]
