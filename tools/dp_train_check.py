#!/usr/bin/env python
"""2+-rank check of `train()` under data parallelism: every rank calls the reference-shaped `train(slices, args)` inside
an initialised NCCL process group; the global batch is split across ranks, the replicas must end with identical
parameters, the data term must fall and the reconstruction must match the phantom as well as the single-GPU run of
tests/test_gpu_e2e.py does.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/dp_train_check.py
"""
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    import nesvor_b200 as nb
    import psnr_phantom as pp
    from nesvor_b200.data.phantom import simulate_slices
    from nesvor_b200.nesvor.sample import sample_points

    args = pp.make_args(dev, n_iter=1500, batch_size=2048, n_samples=64, mask_threshold=0.1, no_loss_sync=True,
                        n_levels_bias=4, no_transformation_optimization=False, output_resolution=1.0,
                        inference_batch_size=1 << 14, n_inference_samples=128, no_output_psf=True)
    torch.manual_seed(rank)  # deliberately different: train() must make the replicas consistent by itself
    slices, volume, _ = simulate_slices(device=dev, n=48, n_stacks=3, res_r=1.0, res_s=1.0, gap=2.0)
    inr, out_slices, mask = nb.train(slices, args)
    flat = torch.cat([p.detach().reshape(-1).float() for p in inr.parameters()])
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(ref, flat))
    grid = pp.phantom_grid(48, 1.0).to(dev)
    gt = volume[0, 0].reshape(-1)
    rec = sample_points(inr, grid, copy.copy(args))
    p_in = pp.psnr(rec.cpu(), gt.cpu(), (gt > 0).cpu())
    out = dict(rank=rank, world=world, replicas_identical=same, psnr_inside=p_in, n_out_slices=len(out_slices),
               mask_fraction=float(mask.mask.float().mean()))
    print(json.dumps(out), flush=True)
    ok = same and p_in > 12.0 and len(out_slices) == len(slices)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
