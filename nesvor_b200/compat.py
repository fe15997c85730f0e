"""Zero-edit drop-in: run the UNMODIFIED reference package (daviddmc/NeSVoR) on libnesvor_b200.

The reference reaches native code through exactly three imports:

* `import nesvor.slice_acq_cuda`          (nesvor/slice_acquisition/slice_acq.py:5-19; else it JIT-compiles its .cu files)
* `import nesvor.transform_convert_cuda`  (nesvor/transform/transform_convert.py:3-18; same fallback)
* `import tinycudann as tcnn`             (nesvor/nesvor/models.py:7; `tcnn.Encoding` :25, `tcnn.Network` :31)

`install()` registers modules under those three names whose functions / classes have the pybind / tcnn signatures and
call the C ABI (include/nesvor_b200.h) -- `forward / backward / adjoint_forward / adjoint_backward`,
`axisangle2mat_{forward,backward} / mat2axisangle_{forward,backward}` (each returning a list of tensors, absent masks as
empty tensors, like slice_acq_cuda.cpp:61-161 and transform_convert_cuda.cpp:27-69), `Encoding(n_input_dims,
encoding_config, dtype)` and `Network(n_input_dims, n_output_dims, network_config)`.  After

    import nesvor_b200.compat as compat; compat.install()
    import nesvor                                   # the reference, untouched

the reference's own `INR`, `NeSVoR`, `train`, `slice_acquisition`, `RigidTransform`, SRR ... run on the B200 kernels
(the unfused path: one native op per reference op; the fused iteration is `nesvor_b200.train(..., args.fused=True)`).
A real `tinycudann`, if installed, is left alone unless `tcnn="force"`.
"""
import importlib
import importlib.util
import sys
import types

import torch

_NAMES = ("nesvor.slice_acq_cuda", "nesvor.transform_convert_cuda", "tinycudann")


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


def install(tcnn: str = "auto") -> dict:
    """Registers the three stand-in modules in `sys.modules`; returns {name: module} of what was installed.
    tcnn = "auto": only when no real `tinycudann` is importable; "force": always; "never": leave `tinycudann` alone."""
    done = {}
    for name, make in (("nesvor.slice_acq_cuda", _slice_acq_module), ("nesvor.transform_convert_cuda", _transform_convert_module)):
        sys.modules[name] = done[name] = make()
    have_real = False
    if tcnn == "auto":
        cur = sys.modules.get("tinycudann")
        have_real = (cur is not None and not getattr(cur, "__nesvor_b200_shim__", False)) or (
            cur is None and importlib.util.find_spec("tinycudann") is not None)
    if tcnn == "force" or (tcnn == "auto" and not have_real):
        sys.modules["tinycudann"] = done["tinycudann"] = _tcnn_module()
    parent = sys.modules.get("nesvor")  # `import nesvor.x as y` resolves through the parent's attribute first
    if parent is not None:
        for name in ("slice_acq_cuda", "transform_convert_cuda"):
            setattr(parent, name, sys.modules["nesvor." + name])
    return done


def uninstall() -> None:
    for name in _NAMES:
        mod = sys.modules.get(name)
        if mod is not None and (name != "tinycudann" or getattr(mod, "__nesvor_b200_shim__", False)):
            del sys.modules[name]
