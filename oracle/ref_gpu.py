"""TEST INFRASTRUCTURE ONLY.  Loader of the reference's own slice-acquisition CUDA extension built for sm_100a by
oracle/build_ref_gpu.sh (oracle/_ref/nesvor_ref_slice_acq_cuda.so): the GPU-side cross-check of kernel B against the
real reference kernels (nesvor/slice_acquisition/slice_acq_cuda.cpp:61-161) and the "reference" arm of its timing.
Returns None when the prebuilt file is absent or cannot be loaded (e.g. a different torch build)."""
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "nesvor_ref_slice_acq_cuda.so")
_mod = None
_err = None


def load():
    global _mod, _err
    if _mod is None and _err is None:
        try:
            import torch  # noqa: F401  (libtorch must be resident before the extension is dlopen'ed)

            if not os.path.exists(SO):
                raise FileNotFoundError(SO)
            spec = importlib.util.spec_from_file_location("nesvor_ref_slice_acq_cuda", SO)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            for name in ("forward", "backward", "adjoint_forward", "adjoint_backward"):
                getattr(mod, name)
            _mod = mod
        except Exception as e:  # absent / ABI mismatch: the cross-check is skipped, never faked
            _err = f"{type(e).__name__}: {e}"
    return _mod


def why_not() -> str:
    return _err or ""
