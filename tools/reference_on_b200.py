#!/usr/bin/env python
"""The UNMODIFIED reference package running on libnesvor_b200 on the GPU (zero-edit drop-in, nesvor_b200/compat.py).

baseline/_ref/nesvor holds the reference's pure-Python layer exactly as it lies under /root/reference (copied by
oracle.build.install_reference_package; git-ignored, travels to the GPU box).  With the three stand-in modules installed,
this script builds the reference's own `NeSVoR` (all heads, bias field and pose optimisation on), loads it with the
parameters of this package's mirror, and on the same batch (same torch RNG state -> same PSF noise) compares

  * the reference's `NeSVoR.forward` losses and parameter gradients  vs  this package's mirror, both under fp16 autocast like
    the reference's loop (must agree to fp round-off: the same op sequence over the same native ops);
  * the reference's `slice_acquisition` / `slice_acquisition_adjoint` / `axisangle2mat` / `mat2axisangle` autograd
    wrappers (its own Function classes, our native modules underneath)  vs  this package's;
  * a few optimiser iterations of the reference's loop body (`train.py:183-197`: autocast forward, GradScaler backward, AdamW
    with the reference's two parameter groups) on the reference's model, to show it trains -- timed, as the "reference code on B200
    kernels" figure.

Prints ONE JSON line ({"available": false, "why": ...} if the reference copy is absent).  Run as a subprocess by
tests/test_gpu_zz_reference.py so that nothing of the reference package enters the test process.
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_PARENT = os.path.join(ROOT, "baseline", "_ref")


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main():
    import torch

    if not os.path.isdir(os.path.join(REF_PARENT, "nesvor")) or not torch.cuda.is_available():
        print(json.dumps({"available": False, "why": "baseline/_ref/nesvor absent" if torch.cuda.is_available() else "no CUDA device"}))
        return
    sys.path.insert(0, REF_PARENT)
    import nesvor_b200.compat as compat

    compat.install()
    from argparse import Namespace

    import nesvor  # noqa: F401  the reference
    import nesvor.nesvor.models as rm
    import nesvor.slice_acquisition as rsa
    import nesvor.transform as rt
    from nesvor.nesvor.train import Dataset as RefDataset  # noqa: F401  (imports the reference's training module)

    import nesvor_b200 as nb
    from nesvor_b200.nesvor.train import build_optimizer, loss_weights

    dev = torch.device("cuda", 0)
    out = {"available": True, "reference_models_file": rm.__file__}
    args = Namespace(n_features_per_level=2, log2_hashmap_size=19, level_scale=1.3819, coarsest_resolution=16.0, finest_resolution=0.5,
                     n_levels_bias=4, depth=1, width=64, n_features_z=15, n_features_slice=16, no_transformation_optimization=False,
                     no_slice_scale=False, no_pixel_variance=False, no_slice_variance=False, single_precision=False, dtype=torch.float16,
                     image_regularization="edge", delta=0.2, device=dev, n_samples=64, n_levels=None, base_resolution=None,
                     weight_transformation=0.1, weight_bias=100.0, weight_image=2.0, learning_rate=5e-3, batch_size=1024)
    g = torch.Generator().manual_seed(0)
    n_s = 12
    ax = (torch.randn(n_s, 6, generator=g) * torch.tensor([0.2, 0.2, 0.2, 4.0, 4.0, 4.0])).to(dev)
    res = torch.tensor([[1.0, 1.0, 3.0]], device=dev).repeat(n_s, 1)
    bb = torch.tensor([[-40.0, -40.0, -40.0], [40.0, 40.0, 40.0]], device=dev)
    torch.manual_seed(1)
    ours = nb.NeSVoR(nb.RigidTransform(ax, True), res, 0.7, bb, args)
    ref = rm.NeSVoR(rt.RigidTransform(ax), res, 0.7, bb, args)
    with torch.no_grad():
        ours.inr.encoding.params.copy_((torch.rand(ours.inr.encoding.params.shape, generator=g) - 0.5).to(dev))
        ours.logit_coef.copy_((torch.randn(n_s, generator=g) * 0.3).to(dev))
        ours.axisangle.add_((torch.randn(n_s, 6, generator=g) * 0.01).to(dev))  # transReg != 0
    ref.load_state_dict(ours.state_dict())
    B = args.batch_size
    xyz = ((torch.rand(B, 3, generator=g) - 0.5) * 40).to(dev)
    xyz[:, 2] = 0
    v = torch.rand(B, generator=g).to(dev)
    idx = torch.randint(0, n_s, (B,), generator=g).to(dev)
    wts = loss_weights(args)

    def run(model):
        model.zero_grad()
        torch.manual_seed(123)  # both forwards draw randn(B, S, 3) first: identical PSF noise
        with torch.autocast("cuda", dtype=torch.float16):  # the reference's loop runs its forward under autocast (train.py:183)
            losses = model(xyz, v, idx)
        total = sum(wts[k] * val for k, val in losses.items() if k in wts and wts[k])
        total.backward()
        return {k: float(val) for k, val in losses.items()}, {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}

    l_ref, g_ref = run(ref)
    l_our, g_our = run(ours)
    out["losses_reference_code"], out["losses_this_package"] = l_ref, l_our
    out["loss_keys_equal"] = sorted(l_ref) == sorted(l_our)
    out["loss_abs_diff"] = {k: abs(l_ref[k] - l_our[k]) for k in l_our}
    out["grad_rel_l2"] = {n: rel_l2(g_ref[n], g_our[n]) for n in g_our}
    # ---- the reference's autograd wrappers over our native modules
    vol = torch.rand(1, 1, 24, 24, 24, generator=g).to(dev).requires_grad_(True)
    psf = nb.get_PSF(res_ratio=(1.0, 1.0, 3.0), device=dev)
    tf = rt.RigidTransform(ax[:5].clone()).matrix().contiguous().requires_grad_(True)
    a = rsa.slice_acquisition(tf, vol, None, None, psf, (20, 20), 1.0, False, False)
    b = nb.slice_acquisition(tf.detach(), vol.detach(), None, None, psf, (20, 20), 1.0, False, False)
    ga = torch.autograd.grad(a.sum(), (tf, vol))
    adj_r = rsa.slice_acquisition_adjoint(tf.detach(), psf, a.detach(), None, None, (24, 24, 24), 1.0, False, True)
    adj_o = nb.slice_acquisition_adjoint(tf.detach(), psf, b, None, None, (24, 24, 24), 1.0, False, True)
    out["wrappers"] = {"slice_acquisition": rel_l2(a, b), "adjoint_equalized": rel_l2(adj_r, adj_o), "grad_finite": bool(all(torch.isfinite(t).all() for t in ga)),
                       "axisangle_round_trip": rel_l2(rt.mat2axisangle(rt.axisangle2mat(ax)), ax)}
    # ---- the reference's loop body on the reference's model (train.py:183-197)
    opt = build_optimizer(ref, args)
    scaler = torch.amp.GradScaler("cuda", init_scale=1.0, enabled=True, growth_factor=2.0, backoff_factor=0.5)
    hist = []
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_it = 30
    e0.record()
    for _ in range(n_it):
        with torch.autocast("cuda", dtype=torch.float16):
            losses = ref(xyz, v, idx)
        loss = sum(wts[k] * val for k, val in losses.items() if k in wts and wts[k])
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        opt.zero_grad()
        hist.append((float(losses["MSE"]), float(loss)))
    e1.record()
    torch.cuda.synchronize()
    out["reference_loop"] = {"iterations": n_it, "mse_first": hist[0][0], "mse_last": hist[-1][0], "total_first": hist[0][1], "total_last": hist[-1][1], "ms_per_iteration": e0.elapsed_time(e1) / n_it,
                             "queries_per_iteration": B * args.n_samples}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
