#!/usr/bin/env python
"""M2: reconstruction PSNR on the 3-D Shepp-Logan phantom, fused B200 path vs the CPU oracle.

Both are trained from identical initial parameters on the identical batch and PSF-noise sequence
(drawn on the host), then resampled on the phantom grid (no output PSF); PSNR is computed against
the phantom (data range 1.0) after a least-squares scalar intensity fit, inside `phantom > 0` and
on the full grid.  Target: |PSNR_ours - PSNR_oracle| <= 0.1 dB.

    python tools/psnr_phantom.py [--cfg 1|2s] [--iters 200] [--batch 2048] [--samples 32] [--json out.json]

cfg 1  = BASELINE config 1: 64^3 phantom, 3 stacks (1.5 mm in-plane, 3 mm thick), 2-level hash grid,
         32-wide MLP, 200 iterations (CPU-feasible for the oracle).
cfg 2s = config-2 model (16 levels, T=2^19, 64 x 3 hidden) on the 64^3 phantom with a reduced batch so
         that the oracle finishes in minutes.
The oracle is the checker here (test infrastructure); the product path is FusedTrainer + fused_render.
"""
import argparse
import json
import os
import sys
import time
from argparse import Namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def make_args(device, **kw):
    a = dict(n_features_per_level=2, log2_hashmap_size=19, level_scale=1.3819, coarsest_resolution=16.0, finest_resolution=0.5,
             n_levels_bias=0, depth=1, width=64, n_features_z=15, n_features_slice=16, no_transformation_optimization=True,
             no_slice_scale=False, no_pixel_variance=False, no_slice_variance=False, single_precision=False,
             weight_transformation=0.1, weight_bias=100.0, image_regularization="edge", weight_image=2.0, delta=0.2,
             learning_rate=5e-3, gamma=0.33, milestones=[0.5, 0.75, 0.9], n_iter=200, batch_size=4096, n_samples=16,
             dtype=torch.float16, device=device, n_levels=None, base_resolution=None, seed=0, fused=True, mask_threshold=1.0)
    a.update(kw)
    return Namespace(**a)


CFGS = {
    "1": dict(sim=dict(n=64, n_stacks=3, res_r=1.0, res_s=1.5, gap=3.0),
              args=dict(coarsest_resolution=16.0, finest_resolution=8.0, level_scale=2.0, width=32, depth=1)),
    "2s": dict(sim=dict(n=64, n_stacks=3, res_r=1.0, res_s=1.0, gap=3.0),
               args=dict(n_levels=16, depth=3, width=64, no_pixel_variance=True, no_slice_variance=True)),
    # reference-default heads (sigma_net + slice variance: the learned variance is what weights the data term up against
    # the edge regulariser) on the 128^3 phantom, poses fixed: the reconstruction-quality run, --ours-only
    "3s": dict(sim=dict(n=128, n_stacks=3, res_r=1.0, res_s=1.0, gap=3.0), args=dict()),
    # BASELINE config 3 at reduced size: stacks simulated at per-slice perturbed ("true") poses (rotation-vector offsets
    # U(+-3 deg)^3, translations U(+-1.5 mm)^3), the NOMINAL stack poses handed to training, reference-default heads,
    # joint pose + INR optimisation (models.py:275-278,357-363, train.py:224)
    "3p": dict(sim=dict(n=64, n_stacks=6, res_r=1.0, res_s=1.0, gap=3.0, motion_deg=3.0, motion_mm=1.5),
               args=dict(no_transformation_optimization=False)),
    # BASELINE config 3 in full: 256^3 phantom, 6 stacks with injected motion, reference-default heads, pose optimisation on,
    # n-samples 256, batch 4096, 8000 iterations (--pose --iters 8000 --batch 4096 --samples 256)
    "3": dict(sim=dict(n=256, n_stacks=6, res_r=1.0, res_s=1.0, gap=3.0, motion_deg=3.0, motion_mm=1.5),
              args=dict(no_transformation_optimization=False)),
    # BASELINE config 2 in full (128^3, 16 levels, 64 x 3 hidden, B = 8192, S = 128, 5000 iterations): --ours-only
    "2": dict(sim=dict(n=128, n_stacks=3, res_r=1.0, res_s=1.0, gap=3.0),
              args=dict(n_levels=16, depth=3, width=64, no_pixel_variance=True, no_slice_variance=True)),
}


def oracle_from_model(model, args, n_slices, resolution, emulate_fp16=False):
    """An oracle model holding the native model's parameters (same logic as tests/test_gpu_fused.build_pair)."""
    from oracle import inr_oracle as io

    enc = model.inr.encoding
    cfg = io.INRConfig(
        n_levels=enc.n_levels, base_resolution=enc.base_resolution, level_scale=args.level_scale, log2_hashmap_size=args.log2_hashmap_size,
        width=args.width, depth=args.depth, n_levels_bias=args.n_levels_bias, no_transformation_optimization=args.no_transformation_optimization,
        no_slice_scale=args.no_slice_scale, no_pixel_variance=args.no_pixel_variance, no_slice_variance=args.no_slice_variance,
        image_regularization=args.image_regularization, n_samples=args.n_samples, delta=model.delta,
        weight_transformation=args.weight_transformation, weight_image=args.weight_image, weight_bias=args.weight_bias, emulate_fp16=emulate_fp16, mlp_bias=False)
    ax = model.axisangle.detach().cpu().float()
    om = io.OracleNeSVoR(cfg, n_slices, ax, resolution.detach().cpu().float(), model.inr.bounding_box.detach().cpu().float())
    P = om.P

    def put(name, t):
        P[name] = t.detach().cpu().float().clone().requires_grad_(name in om.trainable)

    put("table", enc.params)
    nets = [("density_net", model.inr.density_net)]
    if hasattr(model, "sigma_net"):
        nets.append(("sigma_net", model.sigma_net))
    if hasattr(model, "b_net"):
        nets.append(("b_net", model.b_net))
    for prefix, net in nets:
        for i, w in enumerate(net.weight_views()):
            put(f"{prefix}.w{i}", w)
    put("slice_embedding", model.slice_embedding.weight)
    for name in ("logit_coef", "log_var_slice"):
        if hasattr(model, name):
            put(name, getattr(model, name))
    put("axisangle", model.axisangle)
    return om


def psnr(pred: torch.Tensor, gt: torch.Tensor, mask=None) -> float:
    if mask is not None:
        pred, gt = pred[mask], gt[mask]
    pred, gt = pred.double(), gt.double()
    a = (pred * gt).sum() / (pred * pred).sum().clamp_min(1e-30)  # c = softmax * n_s only fixes the mean slice scale
    mse = ((a * pred - gt) ** 2).mean()
    return float(10.0 * torch.log10(1.0 / mse))


def phantom_grid(n, res_r):
    ax = (torch.arange(n, dtype=torch.float32) - (n - 1) / 2.0) * res_r
    zz, yy, xx = torch.meshgrid(ax, ax, ax, indexing="ij")
    return torch.stack((xx, yy, zz), -1).reshape(-1, 3)


def _world_matrix(ax: torch.Tensor) -> torch.Tensor:
    """[n,6] axis-angle (trans_first: x_world = R (x + T), transform.py:259-271) -> [n,4,4] slice-to-world matrices, fp64 on the CPU."""
    from scipy.spatial.transform import Rotation

    ax = ax.detach().double().cpu()
    R = torch.from_numpy(Rotation.from_rotvec(ax[:, :3].numpy()).as_matrix())
    M = torch.eye(4, dtype=torch.float64).repeat(ax.shape[0], 1, 1)
    M[:, :3, :3] = R
    M[:, :3, 3] = torch.einsum("nij,nj->ni", R, ax[:, 3:])
    return M


def pose_error(ax_est: torch.Tensor, ax_true: torch.Tensor, weight: torch.Tensor = None) -> dict:
    """Per-slice pose error up to the global rigid gauge (a reconstruction in a rigidly moved frame is as good):
    G_i = M_est_i M_true_i^-1 maps the true world frame to the estimated one; the residual of G_i against the mean G is
    reported as a rotation angle (degrees) and as the displacement of the slice centre (mm), mean and max over slices."""
    Me, Mt = _world_matrix(ax_est), _world_matrix(ax_true)
    G = Me @ torch.linalg.inv(Mt)
    U, _, Vh = torch.linalg.svd(G[:, :3, :3].mean(0))
    Rm = U @ torch.diag(torch.tensor([1.0, 1.0, float(torch.sign(torch.linalg.det(U @ Vh)))], dtype=torch.float64)) @ Vh
    centre = Mt[:, :3, 3]  # true world position of every slice centre
    moved = torch.einsum("nij,nj->ni", G[:, :3, :3], centre) + G[:, :3, 3]
    tm = (moved - centre @ Rm.T).mean(0)
    disp = (moved - (centre @ Rm.T + tm)).norm(dim=1)
    Rres = Rm.T @ G[:, :3, :3]
    ang = torch.rad2deg(torch.acos(((Rres.diagonal(dim1=1, dim2=2).sum(1) - 1) / 2).clamp(-1, 1)))
    out = {"rot_deg_mean": float(ang.mean()), "rot_deg_median": float(ang.median()), "rot_deg_max": float(ang.max()),
           "centre_mm_mean": float(disp.mean()), "centre_mm_median": float(disp.median()), "centre_mm_max": float(disp.max())}
    if weight is not None:  # weighted by the number of pixels a slice contributes (slices through the phantom's caps carry little signal)
        w = weight.double().cpu() / weight.double().sum().cpu()
        out["rot_deg_pixel_weighted"] = float((ang * w).sum())
        out["centre_mm_pixel_weighted"] = float((disp * w).sum())
    return out


def run_pose_recovery(cfg_name="3p", n_iter=3000, batch=4096, n_samples=64, device=None, log=print):
    """BASELINE config 3 as a WORKLOAD on the product path (nesvor_b200.train, fused kernel A with the pose gradient, transReg,
    fused AdamW): pose error against the injected motion before (nominal stack poses) and after training, and the PSNR of the
    reconstruction with and without pose optimisation on the same data."""
    import nesvor_b200 as nb
    from nesvor_b200.data.phantom import simulate_slices, stack_geometry
    from nesvor_b200.nesvor.fused import attach_render_state, fused_render

    device = device or torch.device("cuda", 0)
    c = CFGS[cfg_name]
    sim = c["sim"]
    out = {"cfg": cfg_name, "iters": n_iter, "batch": batch, "n_samples": n_samples, "sim": sim}
    _, n_slice = stack_geometry(sim["n"], sim["res_r"], sim["res_s"], sim["gap"])
    grid = phantom_grid(sim["n"], sim["res_r"])
    for tag, fixed in (("joint_pose_and_inr", False), ("poses_fixed_at_nominal", True)):
        args = make_args(device, n_iter=n_iter, batch_size=batch, n_samples=n_samples, no_loss_sync=True,
                         **dict(c["args"], no_transformation_optimization=fixed))
        torch.manual_seed(0)
        slices, volume, true_ax = simulate_slices(device=device, **sim)
        keep = torch.tensor([s.stack_idx * n_slice + s.slice_idx for s in slices])
        nominal = torch.cat([s.transformation.axisangle() for s in slices])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        inr, out_slices, _ = nb.train(slices, args)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        est = torch.cat([s.transformation.axisangle() for s in out_slices])
        gt = volume[0, 0].reshape(-1).cpu()
        attach_render_state(inr, args)
        rec = torch.cat([fused_render(inr, grid[i : i + (1 << 18)].to(device), None, 0.0, 1).cpu() for i in range(0, grid.shape[0], 1 << 18)])
        n_px = torch.tensor([float(s.mask.sum()) for s in slices])
        out[tag] = {"pose_error_before": pose_error(nominal, true_ax[keep], n_px), "pose_error_after": pose_error(est, true_ax[keep], n_px),
                    "psnr_inside": psnr(rec, gt, gt > 0), "psnr_full": psnr(rec, gt), "train_wall_s": wall, "n_slices": len(slices)}
        log(tag, json.dumps(out[tag]))
    return out


def run_ours_only(cfg_name="2", n_iter=5000, batch=8192, n_samples=128, device=None, log=print):
    """The product path end to end (nesvor_b200.train with the fused kernel, in-kernel Philox noise, device-side
    epoch shuffles), then PSNR of the resampled volume.  No oracle involved: shows what the path reconstructs at the
    BASELINE configuration's full size and how long it takes."""
    import nesvor_b200 as nb
    from nesvor_b200.data.phantom import simulate_slices
    from nesvor_b200.nesvor.fused import attach_render_state, fused_render

    device = device or torch.device("cuda", 0)
    c = CFGS[cfg_name]
    args = make_args(device, n_iter=n_iter, batch_size=batch, n_samples=n_samples, no_loss_sync=True, **c["args"])
    torch.manual_seed(0)
    slices, volume, _ = simulate_slices(device=device, **c["sim"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    inr, _, _ = nb.train(slices, args)
    torch.cuda.synchronize()
    t_train = time.perf_counter() - t0
    grid = phantom_grid(c["sim"]["n"], c["sim"]["res_r"])
    gt = volume[0, 0].reshape(-1).cpu()
    attach_render_state(inr, args)
    ours = torch.cat([fused_render(inr, grid[i : i + (1 << 20)].to(device), None, 0.0, 1).cpu() for i in range(0, grid.shape[0], 1 << 20)])
    inside = gt > 0
    return {"cfg": cfg_name, "iters": n_iter, "batch": batch, "n_samples": n_samples, "n_slices": len(slices),
            "psnr_ours_inside": psnr(ours, gt, inside), "psnr_ours_full": psnr(ours, gt),
            "train_wall_s_incl_dataset_and_mask": t_train, "queries_per_s_wall": n_iter * batch * n_samples / t_train}


def run(cfg_name="1", n_iter=200, batch=2048, n_samples=32, device=None, threads=None, log=print):
    import nesvor_b200 as nb
    from nesvor_b200.data.phantom import simulate_slices
    from nesvor_b200.nesvor.fused import FusedTrainer, attach_render_state, fused_render
    from nesvor_b200.nesvor.train import Dataset
    from oracle import inr_oracle as io

    device = device or torch.device("cuda", 0)
    torch.set_num_threads(threads or os.cpu_count() or 1)
    c = CFGS[cfg_name]
    args = make_args(device, n_iter=n_iter, batch_size=batch, n_samples=n_samples, **c["args"])
    torch.manual_seed(0)
    slices, volume, _ = simulate_slices(device=device, **c["sim"])
    dataset = Dataset(slices, args)
    model = nb.NeSVoR(dataset.transformation, dataset.resolution, dataset.mean, dataset.bounding_box, args)
    n_slices = len(slices)
    om = oracle_from_model(model, args, n_slices, dataset.resolution)
    opt = io.make_optimizer(om, lr=args.learning_rate)
    trainer = FusedTrainer(model, args)

    xyz_t, v_t, idx_t = dataset.xyz.cpu(), dataset.v.cpu(), dataset.slice_idx.cpu()
    P = xyz_t.shape[0]
    g = torch.Generator().manual_seed(1234)
    perm, cursor = torch.randperm(P, generator=g), 0
    milestones = [int(m * n_iter) for m in args.milestones]
    t_cpu = t_gpu = 0.0
    for it in range(1, n_iter + 1):
        if cursor + batch > P:
            perm, cursor = torch.randperm(P, generator=g), 0
        sel = perm[cursor : cursor + batch]
        cursor += batch
        xyz, v, idx = xyz_t[sel], v_t[sel], idx_t[sel]
        noise = torch.randn(batch, n_samples, 3, generator=g)
        t0 = time.perf_counter()
        losses_o = om.forward(xyz, v, idx, noise)
        om.total_loss(losses_o).backward()
        opt.step()
        opt.zero_grad()
        t_cpu += time.perf_counter() - t0
        t0 = time.perf_counter()
        losses = trainer.step(xyz.to(device), v.to(device), idx.to(device), noise.to(device))
        mse = float(losses["MSE"])
        t_gpu += time.perf_counter() - t0
        if it in milestones:
            trainer.decay_lr(args.gamma)
            for grp in opt.param_groups:
                grp["lr"] *= args.gamma
        if it % max(n_iter // 10, 1) == 0 or it == 1:
            log(f"iter {it:5d}: MSE ours {mse:.5e}  oracle {float(losses_o['MSE']):.5e}")
    trainer.sync_to_model()

    # ---- resample both on the phantom grid (voxel centres; volume centre at the world origin) ----
    grid = phantom_grid(c["sim"]["n"], c["sim"]["res_r"])
    gt = volume[0, 0].reshape(-1).cpu()
    attach_render_state(model.inr, args)
    ours = torch.cat([fused_render(model.inr, grid[i : i + (1 << 18)].to(device), None, 0.0, 1).cpu() for i in range(0, grid.shape[0], 1 << 18)])
    with torch.no_grad():
        orc = torch.cat([om.render(grid[i : i + (1 << 16)], None, 0.0) for i in range(0, grid.shape[0], 1 << 16)])
    inside = gt > 0
    out = {
        "cfg": cfg_name, "iters": n_iter, "batch": batch, "n_samples": n_samples, "n_pixels": int(P), "n_slices": n_slices,
        "psnr_ours_inside": psnr(ours, gt, inside), "psnr_oracle_inside": psnr(orc, gt, inside),
        "psnr_ours_full": psnr(ours, gt), "psnr_oracle_full": psnr(orc, gt),
        "rel_l2_ours_vs_oracle_volume": float((ours - orc).norm() / orc.norm()),
        "oracle_s_per_iter": t_cpu / n_iter, "ours_s_per_iter_incl_h2d_and_sync": t_gpu / n_iter,
        "host_threads": torch.get_num_threads(),
    }
    if not args.no_transformation_optimization:  # joint pose + INR: the two optimisers must move the poses alike
        ax_o = om.P["axisangle"].detach()
        ax_n = model.axisangle.detach().cpu()
        ax_0 = dataset.transformation.axisangle().detach().cpu()
        out["pose_update_rel_l2_ours_vs_oracle"] = float(((ax_n - ax_0) - (ax_o - ax_0)).norm() / (ax_o - ax_0).norm().clamp_min(1e-30))
        out["pose_update_norm_oracle"] = float((ax_o - ax_0).norm())
        out["pose_update_cosine_ours_vs_oracle"] = float(torch.nn.functional.cosine_similarity((ax_n - ax_0).flatten(), (ax_o - ax_0).flatten(), dim=0))
    out["abs_diff_inside_db"] = abs(out["psnr_ours_inside"] - out["psnr_oracle_inside"])
    out["abs_diff_full_db"] = abs(out["psnr_ours_full"] - out["psnr_oracle_full"])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="1", choices=list(CFGS))
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--samples", type=int, default=32, help="PSF samples per pixel (the fused kernel needs 32..256)")
    ap.add_argument("--json", default=None)
    ap.add_argument("--pose", action="store_true", help="config 3 as a workload: pose error before / after joint optimisation + PSNR (product path only)")
    ap.add_argument("--ours-only", action="store_true", help="train with nesvor_b200.train end to end (no oracle) and report PSNR + wall time")
    a = ap.parse_args()
    from nesvor_b200.csrc import build as nsv_build

    nsv_build.build()
    if a.pose:
        out = run_pose_recovery(a.cfg, a.iters, a.batch, a.samples)
    else:
        out = run_ours_only(a.cfg, a.iters, a.batch, a.samples) if a.ours_only else run(a.cfg, a.iters, a.batch, a.samples)
    print(json.dumps(out))
    if a.json:
        with open(a.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
