#!/usr/bin/env python
"""Kernel B against the reference's OWN CUDA extension on the same GPU (test infrastructure + baseline timing).

oracle/build_ref_gpu.sh compiles nesvor/slice_acquisition/slice_acq_cuda.cpp + slice_acq_cuda_kernel.cu for sm_100a
from where they lie under /root/reference (one-token torch-2.x fix applied in a scratch copy) into
oracle/_ref/nesvor_ref_slice_acq_cuda.so, which travels to the GPU box.  This script

  * runs all four operators (forward, backward, adjoint forward with / without `equalize`, adjoint backward) of both
    implementations on identical seeded inputs, with and without masks, and reports relative L2 differences; the same for
    the four pose converters (512 random well-conditioned rows) against the reference's transform_convert_cuda when present;
  * with --reps > 0 times the forward / adjoint operators of both on the BASELINE config-2 stack simulation (231 slices
    of 225^2 pixels, 128^3 volume, PSF of ratio (1, 1, 3)), CUDA events, L2 flushed between launches.

Prints ONE JSON line; {"available": false, "why": ...} when the extension is absent or unusable.  It is run as a
subprocess by tests/test_gpu_zz_reference.py and bench.py so that foreign kernels cannot touch their CUDA contexts.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _lib_mode():
    from nesvor_b200 import _lib

    return int(_lib.lib().nsv_get_slice_acq_exact())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--exact", type=int, default=-1, help="1: the bit-exact (-fmad=false) kernels, 0: the fast product kernels, -1: library default")
    a = ap.parse_args()
    import torch

    from oracle import ref_gpu

    ref = ref_gpu.load()
    if ref is None or not torch.cuda.is_available():
        print(json.dumps({"available": False, "why": ref_gpu.why_not() or "no CUDA device"}))
        return
    import importlib

    from helpers import cuda, slice_acq_case

    sa = importlib.import_module("nesvor_b200.slice_acquisition.slice_acq")
    if a.exact >= 0:
        from nesvor_b200 import _lib

        if hasattr(_lib.lib(), "nsv_set_slice_acq_exact"):
            _lib.lib().nsv_set_slice_acq_exact(int(a.exact))
    empty = torch.empty(0, device="cuda")
    out = {"available": True, "rel_l2": {}}
    for masks in (False, True):
        c = slice_acq_case(seed=5, masks=masks, D=40, H=44, W=48, n=9, h=36, w=34)
        tf, vol, psf = cuda(c["transforms"]), cuda(c["vol"]), cuda(c["psf"])
        vm, sm = (cuda(c["vol_mask"]), cuda(c["slices_mask"])) if masks else (empty, empty)
        ovm, osm = (vm, sm) if masks else (None, None)
        slices, gs, gv = cuda(c["slices"]), cuda(c["grad_slices"]), cuda(c["grad_vol"])
        res = float(c["res_slice"])
        r, o = {}, {}
        r["slices"], r["weight"] = ref.forward(tf, vol, vm, sm, psf, list(c["slice_shape"]), res, True, False)
        o["slices"], o["weight"] = sa.forward(tf, vol, ovm, osm, psf, c["slice_shape"], res, True, False)
        r["bwd_grad_vol"], r["bwd_grad_tf"] = ref.backward(tf, vol, vm, psf, gs, sm, res, False, True, True)
        o["bwd_grad_vol"], o["bwd_grad_tf"] = sa.backward(tf, vol, ovm, psf, gs, osm, res, False, True, True)
        for eq in (0, 1):
            ro = ref.adjoint_forward(tf, psf, slices, sm, vm, list(c["vol_shape"]), res, False, bool(eq))
            ov, ovw = sa.adjoint_forward(tf, psf, slices, osm, ovm, c["vol_shape"], res, False, eq)
            r[f"adj{eq}_vol"], o[f"adj{eq}_vol"] = ro[0], ov
            rvw = ro[1] if eq else empty
            r[f"adjbwd{eq}_grad_slices"], r[f"adjbwd{eq}_grad_tf"] = ref.adjoint_backward(
                tf, gv.clone(), rvw, vm, psf, slices, sm, ro[0] if eq else empty, res, False, bool(eq), True, True)
            o[f"adjbwd{eq}_grad_slices"], o[f"adjbwd{eq}_grad_tf"] = sa.adjoint_backward(
                tf, gv.clone(), ovw, ovm, psf, slices, osm, ov, res, False, eq, True, True)
        # fp64 adjudication: the same operators through this library's _f64 entry points on the double-cast inputs are
        # the common yardstick -- err(ours, f64) and err(reference, f64) say which fp32 implementation is closer to the truth
        d = lambda t: t.double() if t is not None and t.numel() and t.is_floating_point() else t
        t64 = {}
        t64["slices"], t64["weight"] = sa.forward(d(tf), d(vol), ovm, osm, d(psf), c["slice_shape"], res, True, False)
        t64["bwd_grad_vol"], t64["bwd_grad_tf"] = sa.backward(d(tf), d(vol), ovm, d(psf), d(gs), osm, res, False, True, True)
        for eq in (0, 1):
            tv, tvw = sa.adjoint_forward(d(tf), d(psf), d(slices), osm, ovm, c["vol_shape"], res, False, eq)
            t64[f"adj{eq}_vol"] = tv
            t64[f"adjbwd{eq}_grad_slices"], t64[f"adjbwd{eq}_grad_tf"] = sa.adjoint_backward(
                d(tf), d(gv).clone(), tvw, ovm, d(psf), d(slices), osm, tv, res, False, eq, True, True)
        torch.cuda.synchronize()
        tag = "masked" if masks else "plain"
        out["rel_l2"][tag] = {k: rel_l2(o[k], r[k]) for k in r}
        out.setdefault("err_vs_f64", {})[tag] = {k: {"ours": rel_l2(o[k], t64[k]), "reference": rel_l2(r[k], t64[k])} for k in r}
    # ---- pose converters vs the reference's transform_convert_cuda (all four functions)
    tc = ref_gpu.load_transform()
    if tc is not None:
        import ctypes

        from nesvor_b200 import _lib

        # well-conditioned rotations (|w| ~ 0.9 rad, far from pi, where d(axis-angle)/dR amplifies the FMA / non-FMA
        # difference of the two builds); the reference's 11 hand-picked vectors incl. pi - 0.01 are covered bit for bit
        # against the CPU build of the same kernel bodies (tests/test_gpu_pose.py)
        g = torch.Generator().manual_seed(7)
        ax = (torch.randn(512, 6, generator=g) * torch.tensor([0.5, 0.5, 0.5, 30.0, 30.0, 30.0])).cuda()
        n = ax.shape[0]
        gm = torch.randn(n, 3, 4, generator=g).cuda()
        ga = torch.randn(n, 6, generator=g).cuda()
        r_mat = tc.axisangle2mat_forward(ax)[0]
        r = {"a2m_fwd": r_mat, "a2m_bwd": tc.axisangle2mat_backward(gm, ax)[0], "m2a_fwd": tc.mat2axisangle_forward(r_mat)[0],
             "m2a_bwd": tc.mat2axisangle_backward(r_mat, ga)[0]}
        o = {"a2m_fwd": torch.empty(n, 3, 4, device="cuda"), "a2m_bwd": torch.empty(n, 6, device="cuda"), "m2a_fwd": torch.empty(n, 6, device="cuda"),
             "m2a_bwd": torch.empty(n, 3, 4, device="cuda")}
        L, st = _lib.lib(), _lib.stream(ax.device)
        _lib.check(L.nsv_axisangle2mat_fwd_f32(_lib.ptr(ax), _lib.ptr(o["a2m_fwd"]), ctypes.c_int(n), st), "a2m_fwd")
        _lib.check(L.nsv_axisangle2mat_bwd_f32(_lib.ptr(gm), _lib.ptr(ax), _lib.ptr(o["a2m_bwd"]), ctypes.c_int(n), st), "a2m_bwd")
        _lib.check(L.nsv_mat2axisangle_fwd_f32(_lib.ptr(r_mat), _lib.ptr(o["m2a_fwd"]), ctypes.c_int(n), st), "m2a_fwd")
        _lib.check(L.nsv_mat2axisangle_bwd_f32(_lib.ptr(r_mat), _lib.ptr(ga), _lib.ptr(o["m2a_bwd"]), ctypes.c_int(n), st), "m2a_bwd")
        torch.cuda.synchronize()
        out["rel_l2"]["pose_converters"] = {k: rel_l2(o[k], r[k]) for k in r}
    if a.reps > 0:
        from nesvor_b200.data.phantom import STACK_ORIENTATIONS, phantom3d, stack_axisangles, stack_geometry
        from nesvor_b200.transform import RigidTransform, mat_update_resolution
        from nesvor_b200.utils import get_PSF

        dev = torch.device("cuda", 0)
        n = 128
        ss, n_slice = stack_geometry(n, 1.0, 1.0, 3.0)
        vol = torch.tensor(phantom3d(n), dtype=torch.float32, device=dev)[None, None]
        psf = get_PSF(res_ratio=(1.0, 1.0, 3.0), device=dev)
        ax = stack_axisangles(STACK_ORIENTATIONS[:3], n_slice, 3.0).to(dev)
        mat = mat_update_resolution(RigidTransform(ax, trans_first=True).matrix(), 1, 1.0).contiguous()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def timed(fn):
            durs = []
            for i in range(2 + a.reps):
                flush.zero_()
                k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                k0.record()
                res_ = fn()
                k1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    durs.append(k0.elapsed_time(k1))
            return res_, sorted(durs)[len(durs) // 2]  # median: a stray allocator / clock hiccup must not decide a ratio

        s_ref, t_ref_f = timed(lambda: ref.forward(mat, vol, empty, empty, psf, [ss, ss], 1.0, False, False)[0])
        s_our, t_our_f = timed(lambda: sa.forward(mat, vol, None, None, psf, (ss, ss), 1.0, False, False)[0])
        _, t_ref_a = timed(lambda: ref.adjoint_forward(mat, psf, s_ref, empty, empty, [n, n, n], 1.0, False, False)[0])
        _, t_our_a = timed(lambda: sa.adjoint_forward(mat, psf, s_our, None, None, (n, n, n), 1.0, False, False)[0])
        # equalised adjoint (the SRR / PSFreconstruction call, svort/srr.py:118) writes vol and vol_weight
        (v_ref, vw_ref), t_ref_ae = timed(lambda: ref.adjoint_forward(mat, psf, s_ref, empty, empty, [n, n, n], 1.0, False, True)[:2])
        (v_our, vw_our), t_our_ae = timed(lambda: sa.adjoint_forward(mat, psf, s_our, None, None, (n, n, n), 1.0, False, True)[:2])
        g = torch.Generator().manual_seed(11)
        gs_full = torch.randn(s_ref.shape, generator=g).to(dev) * (s_ref > 0)  # cotangent on the acquired pixels
        gv_full = torch.randn(vol.shape, generator=g).to(dev)
        rb, t_ref_b = timed(lambda: ref.backward(mat, vol, empty, psf, gs_full, empty, 1.0, False, True, True))
        ob, t_our_b = timed(lambda: sa.backward(mat, vol, None, psf, gs_full, None, 1.0, False, True, True))
        rab, t_ref_ab = timed(lambda: ref.adjoint_backward(mat, gv_full, empty, empty, psf, s_ref, empty, empty, 1.0, False, False, True, True))
        oab, t_our_ab = timed(lambda: sa.adjoint_backward(mat, gv_full, None, None, psf, s_our, None, None, 1.0, False, 0, True, True))
        pair = lambda a_, b_: {"reference_cuda_extension": a_, "nesvor_b200": b_, "speedup": a_ / b_}
        out["config2_stack_simulation_ms"] = {"slices": int(s_ref.shape[0]), "slice_shape": [ss, ss], "psf_taps": int((psf != 0).sum()),
                                              "forward": pair(t_ref_f, t_our_f), "backward": pair(t_ref_b, t_our_b),
                                              "adjoint_forward": pair(t_ref_a, t_our_a), "adjoint_forward_equalize": pair(t_ref_ae, t_our_ae),
                                              "adjoint_backward": pair(t_ref_ab, t_our_ab),
                                              "rel_l2_ours_vs_reference": {"forward": rel_l2(s_our, s_ref), "adjoint_equalized_vol": rel_l2(v_our, v_ref),
                                                                           "backward_grad_vol": rel_l2(ob[0], rb[0]), "backward_grad_tf": rel_l2(ob[1], rb[1]),
                                                                           "adjoint_backward_grad_slices": rel_l2(oab[0], rab[0]),
                                                                           "adjoint_backward_grad_tf": rel_l2(oab[1], rab[1])}}
        try:
            out["slice_acq_mode"] = "exact" if _lib_mode() else "fast"
        except Exception:
            pass
    print(json.dumps(out))


if __name__ == "__main__":
    main()
