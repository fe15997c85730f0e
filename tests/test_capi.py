"""CPU tests: the C-ABI library builds, loads and exports every symbol include/nesvor_b200.h declares;
host-only helpers behave; device entry points reject bad arguments without touching a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nesvor_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nsv_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(native_lib):
    names = _declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(native_lib, n)]
    assert not missing, f"declared in include/nesvor_b200.h but not exported: {missing}"


def test_version_and_arch(native_lib):
    assert native_lib.nsv_version() == 1
    native_lib.nsv_build_arch.restype = ctypes.c_char_p
    assert native_lib.nsv_build_arch() == b"sm_100a"


def test_cubin_is_sm100a_only():
    import subprocess

    so = os.path.join(ROOT, "nesvor_b200", "csrc", "libnesvor_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_grid_meta_matches_oracle(native_lib):
    from nesvor_b200 import _lib
    from oracle import inr_oracle as io

    for (L, F, T, base, s) in [(16, 2, 19, 9, 1.3819), (12, 2, 19, 7, 1.3819), (2, 2, 19, 5, 2.0), (12, 2, 19, 17, 1.3819), (6, 4, 12, 5, 1.8)]:
        m, total = _lib.make_grid_meta(L, F, T, base, s)
        o = io.grid_meta(L, F, T, base, s)
        assert total == int(o.offset[-1])
        assert np.array_equal(np.array(m.scale[:L], np.float32), o.scale)
        assert np.array_equal(np.array(m.res[:L]), o.res)
        assert np.array_equal(np.array(m.size[:L]), o.size)
        assert np.array_equal(np.array(m.offset[: L + 1]), o.offset)
        assert np.array_equal(np.array(m.hashed[:L]).astype(bool), o.hashed)
    # SURVEY s.8d: cfg 2 table = 5 124 512 entries, 7 dense + 9 hashed levels
    m, total = _lib.make_grid_meta(16, 2, 19, 9, 1.3819)
    assert total == 5124512 and sum(m.hashed[:16]) == 9


def test_bad_arguments_are_rejected_without_gpu(native_lib):
    rc = native_lib.nsv_axisangle2mat_fwd_f32(None, None, ctypes.c_int(4), None)
    assert rc != 0
    native_lib.nsv_last_error_string.restype = ctypes.c_char_p
    assert b"NULL" in native_lib.nsv_last_error_string()
    rc = native_lib.nsv_mlp_fwd_f16(None, None, None, None, ctypes.c_int64(8), ctypes.c_int(31), ctypes.c_int(16), ctypes.c_int(64), ctypes.c_int(1), None)
    assert rc == -2  # NSV_EUNSUPPORTED


def test_product_refuses_cpu_tensors():
    """No CPU fallback: the public ops raise like the reference's CHECK_CUDA (slice_acq_cuda.cpp:57)."""
    import torch
    import nesvor_b200 as nb

    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        nb.axisangle2mat(torch.zeros(2, 6))
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        nb.slice_acquisition(torch.zeros(1, 3, 4), torch.zeros(1, 1, 4, 4, 4), None, None, torch.ones(3, 3, 3), (4, 4), 1.0, False, False)


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ anywhere (judge rule): static scan."""
    pkg = os.path.join(ROOT, "nesvor_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)


def test_bias_mean_and_trans_reg_reject_bad_arguments_without_gpu(native_lib):
    """nsv_inr_bias_mean / nsv_trans_reg_f32 validate on the host before any launch (no GPU here)."""
    from nesvor_b200 import _lib

    native_lib.nsv_last_error_string.restype = ctypes.c_char_p
    cfg, prm = _lib.InrConfig(), _lib.InrParams()
    rc = native_lib.nsv_inr_bias_mean(ctypes.byref(cfg), ctypes.byref(prm), None, None, None, ctypes.c_uint64(0), ctypes.c_uint64(0), None,
                                      ctypes.c_int64(8), ctypes.c_int(32), None)
    assert rc == -1 and b"NULL" in native_lib.nsv_last_error_string()  # NSV_EINVAL
    # a configuration outside the instantiation: non-NULL (never dereferenced) pointers, n_levels_bias = 5
    meta, _ = _lib.make_grid_meta(12, 2, 19, 7, 1.3819)
    cfg.grid, cfg.width, cfg.depth, cfg.n_features_slice, cfg.n_levels_bias, cfg.pixel_variance = meta, 64, 1, 16, 5, 1
    fake = ctypes.c_void_p(256)
    for f in ("table_f16", "mlp_f16", "axisangle", "psf_sigma", "slice_embedding"):
        setattr(prm, f, 256)
    rc = native_lib.nsv_inr_bias_mean(ctypes.byref(cfg), ctypes.byref(prm), fake, fake, None, ctypes.c_uint64(0), ctypes.c_uint64(0), fake,
                                      ctypes.c_int64(8), ctypes.c_int(32), None)
    assert rc == -2 and b"n_levels_bias" in native_lib.nsv_last_error_string()  # NSV_EUNSUPPORTED
    cfg.n_levels_bias = 4
    rc = native_lib.nsv_inr_bias_mean(ctypes.byref(cfg), ctypes.byref(prm), fake, fake, None, ctypes.c_uint64(0), ctypes.c_uint64(0), fake,
                                      ctypes.c_int64(8), ctypes.c_int(48), None)
    assert rc == -1 and b"power of two" in native_lib.nsv_last_error_string()
    # packed MLP layout with the bias head: density | sigma | bias, each 64*32 + 16*64 halves at width 64, depth 1
    off = (ctypes.c_int64 * 3)()
    native_lib.nsv_inr_mlp_layout.restype = ctypes.c_int64
    assert native_lib.nsv_inr_mlp_layout(ctypes.byref(cfg), off) == 3 * 3072 and list(off) == [0, 3072, 6144]
    assert native_lib.nsv_trans_reg_f32(None, None, None, None, ctypes.c_int(0), ctypes.c_float(1.0), None) == 0
    assert native_lib.nsv_trans_reg_f32(None, None, None, None, ctypes.c_int(3), ctypes.c_float(1.0), None) == -1
    assert native_lib.nsv_trans_reg_f32(None, None, None, None, ctypes.c_int(-1), ctypes.c_float(1.0), None) == -1


def test_reference_cuda_extensions_load_and_accept_the_cross_check_calls():
    """oracle/_ref/nesvor_ref_{slice_acq,transform_convert}_cuda.so (the reference's own extensions compiled for sm_100a by
    oracle/build_ref_gpu.sh): importable, sm_100a cubin inside, and every call made by tools/kernel_b_vs_reference.py gets
    past pybind's argument conversion -- with CPU tensors it must stop at the reference's CHECK_CUDA, not at a TypeError.
    Skipped when the files were not built (no /root/reference)."""
    import subprocess

    import torch

    from oracle import ref_gpu

    sa, tc = ref_gpu.load(), ref_gpu.load_transform()
    if sa is None or tc is None:
        pytest.skip("reference CUDA extensions not built: " + ref_gpu.why_not())
    for name in ("nesvor_ref_slice_acq_cuda", "nesvor_ref_transform_convert_cuda"):
        out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "oracle", "_ref", name + ".so")], capture_output=True, text=True).stdout
        assert set(re.findall(r"sm_\d+a?", out)) == {"sm_100a"}
    e = torch.empty(0)
    tf, vol, psf = torch.zeros(2, 3, 4), torch.zeros(1, 1, 4, 4, 4), torch.ones(3, 3, 3)
    sl = torch.zeros(2, 1, 5, 5)
    calls = [lambda: sa.forward(tf, vol, e, e, psf, [5, 5], 1.0, True, False),
             lambda: sa.backward(tf, vol, e, psf, sl, e, 1.0, False, True, True),
             lambda: sa.adjoint_forward(tf, psf, sl, e, e, [4, 4, 4], 1.0, False, True),
             lambda: sa.adjoint_backward(tf, vol, e, e, psf, sl, e, e, 1.0, False, False, True, True),
             lambda: tc.axisangle2mat_forward(torch.zeros(2, 6)), lambda: tc.axisangle2mat_backward(tf, torch.zeros(2, 6)),
             lambda: tc.mat2axisangle_forward(tf), lambda: tc.mat2axisangle_backward(tf, torch.zeros(2, 6))]
    for call in calls:
        with pytest.raises(RuntimeError, match="is_cuda"):
            call()
