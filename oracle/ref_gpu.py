"""TEST INFRASTRUCTURE ONLY.  Loaders of the reference's own CUDA extensions built for sm_100a by oracle/build_ref_gpu.sh
(oracle/_ref/nesvor_ref_slice_acq_cuda.so, oracle/_ref/nesvor_ref_transform_convert_cuda.so): the GPU-side cross-check of
kernel B and the pose converters against the real reference kernels (nesvor/slice_acquisition/slice_acq_cuda.cpp:61-161,
nesvor/transform/transform_convert_cuda.cpp:27-69) and the "reference" arm of kernel B's timing.  `load*()` return None when
the prebuilt file is absent or cannot be loaded (e.g. a different torch build); `why_not()` says why."""
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_MODULES = {
    "nesvor_ref_slice_acq_cuda": ("forward", "backward", "adjoint_forward", "adjoint_backward"),
    "nesvor_ref_transform_convert_cuda": ("axisangle2mat_forward", "axisangle2mat_backward", "mat2axisangle_forward", "mat2axisangle_backward"),
}
SO = os.path.join(HERE, "_ref", "nesvor_ref_slice_acq_cuda.so")
_loaded = {}
_errors = {}


def _load(name):
    if name not in _loaded and name not in _errors:
        try:
            import torch  # noqa: F401  (libtorch must be resident before the extension is dlopen'ed)

            path = os.path.join(HERE, "_ref", name + ".so")
            if not os.path.exists(path):
                raise FileNotFoundError(path)
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            for fn in _MODULES[name]:
                getattr(mod, fn)
            _loaded[name] = mod
        except Exception as e:  # absent / ABI mismatch: the cross-check is skipped, never faked
            _errors[name] = f"{type(e).__name__}: {e}"
    return _loaded.get(name)


def load():
    """The slice-acquisition extension (forward, backward, adjoint_forward, adjoint_backward)."""
    return _load("nesvor_ref_slice_acq_cuda")


def load_transform():
    """The pose-converter extension (axisangle2mat_{forward,backward}, mat2axisangle_{forward,backward})."""
    return _load("nesvor_ref_transform_convert_cuda")


def why_not() -> str:
    return "; ".join(_errors.values())
