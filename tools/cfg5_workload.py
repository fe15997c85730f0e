#!/usr/bin/env python
"""BASELINE config 5 as a WORKLOAD (not a head set): fetal-brain-sized synthetic -- 138^3 phantom at 0.8 mm (~110 mm box), 9 stacks
x 30 slices (0.8 mm in-plane, 3 mm thick), every slice multiplied by a smooth bias field, per-slice motion (2 deg / 1 mm);
reference defaults + `--n-levels-bias 4`, pixel + slice variance on, pose optimisation on, finest resolution 0.5, S = 256,
4096 pixels per rank and iteration (2^20 queries per rank) -- through `nesvor_b200.train()` inside the process group the
launcher set up (data parallel over all ranks; a single process runs the same global batch on one GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 tools/cfg5_workload.py --iters 10000
    python tools/cfg5_workload.py --iters 10000 --ranks-equivalent 4          # the same global batch on one GPU

Prints ONE JSON line on rank 0: ms per iteration (whole train() wall time / iterations), final losses, PSNR of the
reconstruction against the phantom (inside / full grid, least-squares intensity fit), replicas identical.
"""
import argparse
import copy
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10000)
    ap.add_argument("--ranks-equivalent", type=int, default=0, help="single process: use the global batch of this many ranks")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import nesvor_b200 as nb
    import psnr_phantom as pp
    from nesvor_b200.data.phantom import simulate_slices
    from nesvor_b200.nesvor.sample import sample_points

    n, res = 138, 0.8
    eq = max(world, a.ranks_equivalent, 1)
    args = pp.make_args(dev, n_iter=a.iters, batch_size=4096 * eq, n_samples=256, n_levels_bias=4, finest_resolution=0.5,
                        no_transformation_optimization=False, mask_threshold=0.1, no_loss_sync=True, output_resolution=res,
                        inference_batch_size=1 << 15, n_inference_samples=128, no_output_psf=True)
    torch.manual_seed(0)
    slices, volume, true_ax = simulate_slices(device=dev, n=n, n_stacks=9, res_r=res, res_s=res, gap=3.0, n_slice=30, motion_deg=2.0, motion_mm=1.0)
    for s in slices:  # smooth multiplicative bias field in slice coordinates, different per stack
        h, w = s.image.shape[-2:]
        yy, xx = torch.meshgrid(torch.linspace(-1, 1, h, device=dev), torch.linspace(-1, 1, w, device=dev), indexing="ij")
        s.image = s.image * torch.exp(0.3 * torch.cos(1.5 * xx + 0.5 * s.stack_idx) * torch.cos(1.1 * yy))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    inr, out_slices, mask = nb.train(slices, args)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    flat = torch.cat([p.detach().reshape(-1).float() for p in inr.parameters()])
    same = True
    if world > 1:
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(ref, flat))
        t = torch.tensor([1.0 if same else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        same = bool(t.item() == 1.0)
    if rank == 0:
        grid = pp.phantom_grid(n, res).to(dev)
        gt = volume[0, 0].reshape(-1)
        rec = sample_points(inr, grid, copy.copy(args))
        keep = torch.tensor([s.stack_idx * 30 + s.slice_idx for s in slices])
        nominal = torch.cat([s.transformation.axisangle() for s in slices])
        est = torch.cat([s.transformation.axisangle() for s in out_slices])
        out = {"workload": "BASELINE config 5", "n_gpus": world, "global_batch_pixels": args.batch_size, "n_samples": 256, "iterations": a.iters,
               "queries_per_iteration_global": args.batch_size * 256, "n_slices": len(slices), "n_levels": int(inr.encoding.n_levels),
               "train_wall_s_incl_dataset_and_mask": wall, "ms_per_iteration_wall": 1e3 * wall / a.iters,
               "queries_per_s_wall": a.iters * args.batch_size * 256 / wall, "replicas_identical": same, "finite": bool(torch.isfinite(rec).all()),
               "psnr_inside": pp.psnr(rec.cpu(), gt.cpu(), (gt > 0).cpu()), "psnr_full": pp.psnr(rec.cpu(), gt.cpu()),
               "pose_error_before": pp.pose_error(nominal, true_ax[keep]), "pose_error_after": pp.pose_error(est, true_ax[keep])}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
