#!/usr/bin/env python
"""2+-rank check of the data-parallel path with the bias-field head (BASELINE config-5 heads: b_net + pixel / slice
variance + pose optimisation).  biasReg = mean(log_bias)^2 couples all samples of the GLOBAL batch, so the ranks average
their nsv_inr_bias_mean results before kernel A runs.  Checked on every rank:

  * losses[4] after the data-parallel forward/backward == mean(log_bias) of the whole global batch evaluated by ONE
    launch over the concatenated batch (same parameters, same PSF noise);
  * mean over ranks of the per-rank gradients == gradient of the single-launch global batch (rel-L2 <= 2e-3: float-atomic
    ordering + fp16 backward operands with different loss-scale exponents);
  * three `step_distributed` iterations through the peer-memory optimiser (losses live in the symmetric buffer there)
    leave identical fp16 parameters on every rank and finite losses including biasReg.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/dp_bias_check.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def rel_l2(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    import nesvor_b200 as nb
    import psnr_phantom as pp
    from nesvor_b200.data.phantom import simulate_slices
    from nesvor_b200.nesvor.fused import FusedState, FusedTrainer
    from nesvor_b200.nesvor.train import Dataset

    B, S = 512, 64  # per rank
    args = pp.make_args(dev, batch_size=B, n_samples=S, n_levels_bias=4, no_transformation_optimization=False)
    torch.manual_seed(0)
    slices, _, _ = simulate_slices(device=dev, n=48, n_stacks=3, res_r=1.0, res_s=1.0, gap=2.0, motion_deg=2.0, motion_mm=1.0)
    dataset = Dataset(slices, args)
    torch.manual_seed(7)  # identical initial parameters on every rank
    model = nb.NeSVoR(dataset.transformation, dataset.resolution, dataset.mean, dataset.bounding_box, args)
    g = torch.Generator().manual_seed(11)  # identical global batch on every rank
    P = dataset.xyz.shape[0]
    sel = torch.randint(0, P, (B * world,), generator=g).to(dev)
    noise = torch.randn(B * world, S, 3, generator=g).to(dev)
    xyz, v, idx = dataset.xyz[sel], dataset.v[sel], dataset.slice_idx[sel]
    mine = slice(rank * B, (rank + 1) * B)

    # ---- per-rank chunks + averaged mean(log_bias) + averaged gradients ----
    st = FusedState(model.inr, args, model, n_batch_samples=B * S)
    st.grad.zero_()
    losses, _ = st.forward_backward(xyz[mine], v[mine], idx[mine], noise[mine], dist=dist, world=world)
    mean_dp = float(losses[4])
    loss_dp = losses[:4].clone()
    dist.all_reduce(loss_dp)
    loss_dp /= world
    loss_dp[2] = losses[2]  # biasReg is already the global value on every rank
    g_dp = st.grad[: st.n_total].clone()
    dist.all_reduce(g_dp)
    g_dp /= world

    # ---- the same global batch in one launch ----
    st1 = FusedState(model.inr, args, model, n_batch_samples=B * S * world)
    st1.grad.zero_()
    losses1, _ = st1.forward_backward(xyz, v, idx, noise)
    torch.cuda.synchronize()
    mean_1 = float(losses1[4])
    g_1 = st1.grad[: st1.n_total]
    out = dict(rank=rank, world=world, mean_log_bias_dp=mean_dp, mean_log_bias_single=mean_1,
               losses_dp=[float(x) for x in loss_dp], losses_single=[float(x) for x in losses1[:4]],
               grad_rel_l2=rel_l2(g_dp, g_1))
    for name in ("table", "mlp", "slice_embedding", "axisangle"):
        out["grad_rel_l2_" + name] = rel_l2(g_dp[st.offsets[name]], g_1[st1.offsets[name]])
    ok = abs(mean_dp - mean_1) <= 1e-5 * max(1.0, abs(mean_1)) and out["grad_rel_l2"] <= 2e-3
    ok = ok and all(abs(a - b) <= 2e-3 * abs(b) + 1e-6 for a, b in zip(out["losses_dp"], out["losses_single"]))

    # ---- three optimiser iterations through the peer-memory path ----
    trainer = FusedTrainer(model, args)
    gi = torch.Generator().manual_seed(100 + rank)
    last = {}
    for _ in range(3):
        s2 = torch.randint(0, P, (B,), generator=gi).to(dev)
        last = trainer.step_distributed(dist, world, dataset.xyz[s2], dataset.v[s2], dataset.slice_idx[s2])
    torch.cuda.synchronize()
    n = trainer.state.n_train
    ref = trainer.state.flat16[:n].clone()
    dist.broadcast(ref, src=0)
    same = bool((ref == trainer.state.flat16[:n]).all())
    finite = all(bool(torch.isfinite(x)) for x in last.values())
    # batch-independent terms must agree across ranks (they read the per-slice fp32 parameters every rank mirrors)
    tr = torch.stack([last["transReg"].detach().float().reshape(()), last["biasReg"].detach().float().reshape(())])
    tr_all = [torch.empty_like(tr) for _ in range(world)]
    dist.all_gather(tr_all, tr)
    consistent = all(bool(torch.equal(t, tr_all[0])) for t in tr_all)
    out.update(dp_mode=trainer.dp_mode, replicas_identical=same, rank_independent_terms_identical=consistent,
               losses_after_3_steps={k: float(x) for k, x in last.items()})
    ok = ok and same and finite and consistent and "biasReg" in last
    print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
