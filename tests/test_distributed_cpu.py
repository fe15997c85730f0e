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
        out.put((g * scale).numpy().copy())  # by value (a torch tensor would travel as a shared-memory handle of this process)
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
    torch.testing.assert_close(torch.from_numpy(got), ref, rtol=1e-9, atol=1e-12)


def test_shard_bounds_cover_everything():
    from nesvor_b200.nesvor.distributed import shard_bounds

    for n in (0, 1, 7, 8192, 8193):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _bias_worker(rank, world, port, out):
    """Bias-field head under data parallelism, host logic only (the oracle stands in for the kernels): every rank evaluates
    mean(log_bias) of its shard, the ranks average it (FusedState.forward_backward(dist=..., world=...)), back-propagate
    their batch-mean terms + w_bias * 2 * mean_global * (their shard's mean log_bias), and the all-reduced gradient must
    equal the single-process gradient of MSE + logVar + w_bias * biasReg over the whole batch."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nesvor_b200.nesvor.distributed import allreduce_gradient, shard_batch, shard_bounds

    om, batch, noise = _make_bias()
    n = batch["v"].shape[0]
    lo, hi = shard_bounds(n, rank, world)
    sb = shard_batch(batch, rank, world)
    for k in om.trainable:
        om.P[k].grad = None
    losses, aux = om.forward(sb["xyz"], sb["v"], sb["slice_idx"], noise[lo:hi], return_aux=True)
    local_mean = losses["biasReg"].detach().sqrt()  # |mean|; recover the sign from the forward below
    mean_local = _mean_log_bias(om, sb, noise[lo:hi])
    assert abs(float(local_mean) - abs(float(mean_local))) < 1e-12
    m = mean_local.detach().clone().reshape(1)
    dist.all_reduce(m)
    m /= world  # equal shards: the global mean
    # d(w * mean_g^2)/dtheta restricted to this rank's samples, in the rank-mean convention (the all-reduce averages):
    # w * 2 * mean_g * d(mean_local)/dtheta
    (losses["MSE"] + losses["logVar"] + om.cfg.weight_bias * 2.0 * float(m) * mean_local).backward()
    g = torch.cat([(om.P[k].grad if om.P[k].grad is not None else torch.zeros_like(om.P[k])).reshape(-1) for k in om.trainable])
    scale = allreduce_gradient(g, dist, world, local_count=hi - lo, global_count=n)
    if rank == 0:
        out.put(((g * scale).numpy().copy(), float(m)))  # by value: a torch tensor travels as a shared-memory handle that dies with the sender
    dist.barrier()
    dist.destroy_process_group()


def _make_bias():
    from oracle import inr_oracle as io

    cfg = io.INRConfig(n_levels=4, base_resolution=4, level_scale=1.6, log2_hashmap_size=10, width=16, depth=1, n_samples=4,
                       n_levels_bias=2, no_transformation_optimization=False, weight_bias=100.0)
    n_slices = 5
    g = torch.Generator().manual_seed(3)
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


def _mean_log_bias(om, batch, noise):
    """mean(log_bias) of a batch with its graph (what nsv_inr_bias_mean computes forward-only)."""
    from oracle import inr_oracle as io

    cfg, P = om.cfg, om.P
    B, S = batch["xyz"].shape[0], noise.shape[1]
    idx = batch["slice_idx"]
    pts = batch["xyz"][:, None] + noise * om.psf_sigma[idx][:, None]
    mat = io.axisangle2mat(P["axisangle"][idx]).view(B, 1, 3, 4)
    x = io.mat_transform_points(mat, pts, True)
    _, pe, _ = om.inr_forward(x)
    se = P["slice_embedding"][idx][:, None].expand(-1, S, -1).reshape(-1, cfg.n_features_slice)
    lb = om._mlp("b_net", torch.cat([se, pe[..., : cfg.n_levels_bias * cfg.n_features_per_level]], -1))[..., 0]
    return lb.mean()


def test_two_rank_bias_head_gradient_equals_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_bias_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    g_dp, mean_dp = out.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    om, batch, noise = _make_bias()
    for k in om.trainable:
        om.P[k].grad = None
    losses = om.forward(batch["xyz"], batch["v"], batch["slice_idx"], noise)
    (losses["MSE"] + losses["logVar"] + om.cfg.weight_bias * losses["biasReg"]).backward()
    g_ref = torch.cat([(om.P[k].grad if om.P[k].grad is not None else torch.zeros_like(om.P[k])).reshape(-1) for k in om.trainable])
    assert abs(mean_dp**2 - float(losses["biasReg"].detach())) < 1e-12
    torch.testing.assert_close(torch.from_numpy(g_dp), g_ref, rtol=1e-9, atol=1e-12)


def _dataset_worker(rank, world, port, out):
    """Dataset.get_batch under a process group: ranks seeded differently must still walk the same shuffled table."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nesvor_b200.nesvor.distributed import shard_batch
    from nesvor_b200.nesvor.train import Dataset

    ds = Dataset.__new__(Dataset)
    P = 50
    ds.xyz = torch.arange(P * 3, dtype=torch.float32).view(P, 3)
    ds.v = torch.arange(P, dtype=torch.float32)
    ds.slice_idx = torch.arange(P) % 7
    ds.count, ds.epoch, ds.dist = P, 0, dist  # exhausted: the first call reshuffles
    torch.manual_seed(100 + rank)
    seen = []
    for _ in range(7):  # 16-pixel batches from a 50-pixel table: crosses two epoch boundaries
        b = ds.get_batch(16, torch.device("cpu"))
        assert torch.equal(b["xyz"][:, 0] / 3, b["v"]) and torch.equal(b["slice_idx"], b["v"].long() % 7)  # rows stay together
        seen.append((b["v"].tolist(), shard_batch(b, rank, world)["v"].tolist()))  # plain lists: pickled by value
    out.put((rank, seen, ds.epoch))
    dist.barrier()
    dist.destroy_process_group()


def test_dataset_batches_are_identical_across_ranks_and_sharded_disjointly():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_dataset_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict()
    for _ in range(world):
        r, seen, epoch = out.get(timeout=300)
        res[r] = (seen, epoch)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] == 3
    for (g0, s0), (g1, s1) in zip(res[0][0], res[1][0]):
        assert g0 == g1  # the same global batch on both ranks
        assert s0 + s1 == g0  # rank chunks tile it
    first_epoch = [x for g, _ in res[0][0][:3] for x in g]
    assert len(set(first_epoch)) == 48  # 3 batches of one permutation: no repeats
