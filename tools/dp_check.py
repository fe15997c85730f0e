#!/usr/bin/env python
"""2+-rank check of the fused data-parallel optimiser (reduce-scatter + AdamW + all-gather over NVLink peer memory,
nsv_adamw_step_dp) against NCCL all-reduce + nsv_adamw_step: both trainers start from identical parameters and see
identical per-rank batches and PSF noise; after a few iterations the fp16 parameters every rank trains with and the
gathered fp32 master must agree (bit-identical for 2 ranks: a + b has one summation order).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py
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
    from nesvor_b200.nesvor.fused import FusedTrainer
    from nesvor_b200.nesvor.train import Dataset

    args = pp.make_args(dev, batch_size=1024, n_samples=64, no_transformation_optimization=False)  # every head + pose gradient
    torch.manual_seed(0)
    slices, _, _ = simulate_slices(device=dev, n=48, n_stacks=3, res_r=1.0, res_s=1.0, gap=2.0, motion_deg=2.0, motion_mm=1.0)
    dataset = Dataset(slices, args)
    trainers = {}
    # "peer": the fused kernel with the ranks' rendezvous inside it (nsv_adamw_step_dp_sync, the default);
    # "peer_host": the same kernel bracketed by two host-launched symmetric-memory barriers (nsv_adamw_step_dp)
    # "peer_mc": host barriers + the NVSwitch-multicast variant of the kernel (nsv_adamw_step_dp_mc), when the fabric has multicast
    for mode in ("peer", "peer_host", "peer_hybrid", "peer_mc", "allreduce"):
        torch.manual_seed(7)  # identical initial parameters on every rank and for all trainers
        model = nb.NeSVoR(dataset.transformation, dataset.resolution, dataset.mean, dataset.bounding_box, args)
        a = copy.copy(args)
        a.dp_optimizer = "peer" if mode.startswith("peer") else mode
        a.dp_sync = {"peer_host": "host", "peer_hybrid": "hybrid", "peer_mc": "host"}.get(mode, "kernel")  # every synchronisation scheme is exercised whatever the default is
        a.dp_multimem = mode == "peer_mc"
        trainers[mode] = FusedTrainer(model, a)
    g = torch.Generator().manual_seed(100 + rank)
    P = dataset.xyz.shape[0]
    tp, th, ty, ta = trainers["peer"], trainers["peer_host"], trainers["peer_hybrid"], trainers["allreduce"]
    tm = trainers["peer_mc"]
    for name in ("slice_embedding", "logit_coef", "log_var_slice", "axisangle"):
        t = tp.state.seg(name)
        if t is not None:
            setattr(tp, "_initial_" + name, t.clone())
    for it in range(6):
        sel = torch.randint(0, P, (args.batch_size,), generator=g).to(dev)
        noise = torch.randn(args.batch_size, args.n_samples, 3, generator=g).to(dev)
        batch = dict(xyz=dataset.xyz[sel], v=dataset.v[sel], slice_idx=dataset.slice_idx[sel])
        # kernel A accumulates with float atomics (run-to-run round-off that Adam's normalisation amplifies), so the two
        # optimiser paths are compared on the SAME per-rank gradient: one forward / backward, copied into both trainers
        for tr in (tp, th, ty, tm, ta):
            tr.iteration += 1
            tr.state.losses.zero_()
        ta.state.forward_backward(batch["xyz"], batch["v"], batch["slice_idx"], noise)
        if it == 0:
            tp._setup_dp(dist, world)
            th._setup_dp(dist, world)
            ty._setup_dp(dist, world)
            tm._setup_dp(dist, world)
        for tr in (tp, th, ty, tm):
            tr.state.grad[: tr.state.n_total].copy_(ta.state.grad[: ta.state.n_total])
        for tr in (tp, th, ty, tm, ta):
            tr._dp_update(dist, world)
        # the next forward must see the same parameters in both trainers (checked at the end); keep them in lockstep
    torch.cuda.synchronize()
    assert tp.dp_mode == "peer" and ta.dp_mode == "allreduce", (tp.dp_mode, ta.dp_mode)
    n = tp.state.n_train
    assert th.dp_mode == "peer"
    d_host = max((th.state.flat16[:n].float() - tp.state.flat16[:n].float()).abs().max().item(),
                 (ty.state.flat16[:n].float() - tp.state.flat16[:n].float()).abs().max().item())  # the three synchronisation schemes: same kernel, same bits
    d16 = (tp.state.flat16[:n].float() - ta.state.flat16[:n].float()).abs().max().item()
    # multicast variant: the switch sums in its own order (exact for 2 ranks: one addition), every rank must still hold the same bits
    mc_on = bool(getattr(tm, "dp_multimem_active", False))
    d_mc = (tm.state.flat16[:n].float() - ta.state.flat16[:n].float()).abs().max().item()
    d_mc_tail = 0.0
    for name in ("slice_embedding", "logit_coef", "log_var_slice", "axisangle"):
        pa, pb = tm.state.seg(name), ta.state.seg(name)
        if pa is not None and pa.numel():
            d_mc_tail = max(d_mc_tail, (pa - pb).abs().max().item())
    ref_mc = tm.state.flat16[:n].clone()
    dist.broadcast(ref_mc, src=0)
    same_mc = bool((ref_mc == tm.state.flat16[:n]).all())
    changed = (tp.state.flat16[:n] != 0).float().mean().item()
    # the per-slice parameters kernel A reads in fp32 (slice embedding, slice scale / variance, poses) must be current on
    # EVERY rank, not only on the owner of their optimiser shard (fp32 mirror written by nsv_adamw_step_dp)
    d_tail = 0.0
    for name in ("slice_embedding", "logit_coef", "log_var_slice", "axisangle"):
        pa, pb = tp.state.seg(name), ta.state.seg(name)
        if pa is not None and pa.numel():
            d_tail = max(d_tail, (pa - pb).abs().max().item())
            assert (pa != getattr(tp, "_initial_" + name)).any(), name + " did not move"
    tp.sync_to_model()
    ta.sync_to_model()
    d32 = (tp.state.flat[:n] - ta.state.flat[:n]).abs().max().item()
    scale = ta.state.flat[:n].abs().max().item()
    # every rank must hold the same fp16 parameters
    ref = tp.state.flat16[:n].clone()
    dist.broadcast(ref, src=0)
    same = bool((ref == tp.state.flat16[:n]).all())
    # exchange time of each scheme, alone (gradient mean + AdamW + refresh; gradients are zero now, the traffic is the same)
    timing = {}
    for name, tr in (("peer_host", th), ("peer_mc", tm), ("allreduce", ta)):
        for _ in range(3):
            tr.iteration += 1
            tr._dp_update(dist, world)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            tr.iteration += 1
            tr._dp_update(dist, world)
        e1.record()
        torch.cuda.synchronize()
        timing[name] = e0.elapsed_time(e1) / 20
    out = dict(rank=rank, world=world, multicast=mc_on, max_abs_diff_fp16_multicast=d_mc, max_abs_diff_per_slice_fp32_multicast=d_mc_tail,
               replicas_identical_multicast=same_mc, exchange_ms=timing, max_abs_diff_kernel_vs_host_sync=d_host, rendezvous_timeouts=int(tp.state.dp_flags[33]), max_abs_diff_fp16=d16, max_abs_diff_fp32_master=d32, max_abs_diff_per_slice_fp32=d_tail,
               param_scale=scale, replicas_identical=same, nonzero_frac=changed)
    print(json.dumps(out), flush=True)
    ok = same and d_host == 0.0 and int(tp.state.dp_flags[33]) == 0 and d16 <= (0.0 if world == 2 else 2e-3 * scale) and d32 <= (0.0 if world == 2 else 1e-4 * scale)
    ok = ok and d_tail <= (0.0 if world == 2 else 1e-4 * scale)
    ok = ok and same_mc and d_mc <= (0.0 if world == 2 else 2e-3 * scale) and d_mc_tail <= (0.0 if world == 2 else 1e-4 * scale)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
