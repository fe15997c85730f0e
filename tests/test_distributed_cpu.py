"""CPU test of the N>1 path (world_size 2, gloo): sharding a global batch across ranks, computing
per-rank batch-mean gradients (oracle model as the compute stand-in -- the CUDA kernel needs a GPU),
one all-reduce over the flat gradient, and the unscale factor together reproduce the single-process
gradient of the full batch.  This is exactly the logic FusedTrainer.step_distributed runs over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make(seed=0):
    from oracle import inr_oracle as io

    cfg = io.INRConfig(n_levels=4, base_resolution=4, level_scale=1.6, log2_hashmap_size=10, width=16, depth=1, n_samples=4,
                       no_transformation_optimization=True)
    n_slices = 5
    g = torch.Generator().manual_seed(seed)
    ax = torch.randn(n_slices, 6, generator=g) * 0.1
    res = torch.tensor([[1.0, 1.0, 3.0]]).repeat(n_slices, 1)
    bb = torch.tensor([[-20.0] * 3, [20.0] * 3])
    om = io.OracleNeSVoR(cfg, n_slices, ax, res, bb, dtype=torch.float64)
    with torch.no_grad():
        om.P["table"].copy_(torch.randn(om.P["table"].shape, generator=g, dtype=torch.float64) * 0.3)
    B = 12
    batch = {"xyz": (torch.rand(B, 3, generator=g, dtype=torch.float64) - 0.5) * 20, "v": torch.rand(B, generator=g, dtype=torch.float64),
             "slice_idx": torch.randint(0, n_slices, (B,), generator=g)}
    noise = torch.randn(B, cfg.n_samples, 3, generator=g, dtype=torch.float64)
    return om, batch, noise


def _flat_grad(om, batch, noise):
    for k in om.trainable:
        om.P[k].grad = None
    losses = om.forward(batch["xyz"], batch["v"], batch["slice_idx"], noise)
    # batch-mean terms only: the image regulariser's "mean - 1" offset and transReg are batch independent
    (losses["MSE"] + losses["logVar"] + om.cfg.weight_image * losses["imageReg"]).backward()
    return torch.cat([om.P[k].grad.reshape(-1) for k in om.trainable])


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nesvor_b200.nesvor.distributed import allreduce_gradient, shard_batch, shard_bounds

    om, batch, noise = _make()
    n = batch["v"].shape[0]
    lo, hi = shard_bounds(n, rank, world)
    g = _flat_grad(om, shard_batch(batch, rank, world), noise[lo:hi])
    scale = allreduce_gradient(g, dist, world, local_count=hi - lo, global_count=n)
    if rank == 0:
        out.put((g * scale).clone())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = out.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    om, batch, noise = _make()
    ref = _flat_grad(om, batch, noise)
    torch.testing.assert_close(got, ref, rtol=1e-9, atol=1e-12)


def test_shard_bounds_cover_everything():
    from nesvor_b200.nesvor.distributed import shard_bounds

    for n in (0, 1, 7, 8192, 8193):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
