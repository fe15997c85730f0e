#!/usr/bin/env python
"""Kernel-A-only timing sweep over the gather/scatter tuning hooks (BASELINE config 2 by default).

    python tools/bench_kernel_a.py [--variants "agg:fast,agg:fast,..."] [--reps 20] [--cfg 2|3|5]

Prints one line per variant: mean / min kernel time (CUDA events, L2 flushed between launches)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="8192:1,0:1,0:0,8192:0,40000:1,300000:1")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--cfg", type=int, default=2)
    ap.add_argument("--impl", default="auto")
    ap.add_argument("--ablate", default="0", help="comma list of NSV_ABLATE masks (1: no table loads, 2: no table reductions, 4: no MLP chain)")
    ap.add_argument("--groups", default="-1", help="comma list of nsv_set_fused_tc_groups values (2, 3, -1 default)")
    ap.add_argument("--locality", default="0", help="comma list: 0 random order inside the batch + strided tiles, 1 (slice, y, x)-ordered batch + contiguous tiles per CTA")
    ap.add_argument("--timers", action="store_true", help="per-phase warp-cycle breakdown of the tcgen05 kernel (profiling build)")
    a = ap.parse_args()
    import torch

    import bench
    from nesvor_b200 import _lib
    from nesvor_b200.csrc import build as nsv_build

    nsv_build.build()
    import nesvor_b200 as nb
    from nesvor_b200.data.phantom import simulate_slices
    from nesvor_b200.nesvor.fused import FusedTrainer
    from nesvor_b200.nesvor.train import Dataset

    dev = torch.device("cuda", 0)
    _lib.set_fused_impl(a.impl)
    if a.cfg == 2:
        args = bench.make_args(dev)
        n, n_stacks, kw = 128, 3, {}
    elif a.cfg == 5:  # config-5 heads: defaults + bias field on 4 levels (b_net), finest resolution 0.5; 9 stacks x 30 slices at 0.8 mm
        args = bench.make_args(dev, depth=1, no_pixel_variance=False, no_slice_variance=False, no_transformation_optimization=False,
                               n_levels=None, n_samples=256, batch_size=4096, n_levels_bias=4, finest_resolution=0.5)
        n, n_stacks, kw = 138, 9, dict(res_r=0.8, res_s=0.8, n_slice=30)
    else:  # config-3 heads: defaults (depth 1, sigma_net, slice variance, pose optimisation), S = 256
        args = bench.make_args(dev, depth=1, no_pixel_variance=False, no_slice_variance=False, no_transformation_optimization=False,
                               n_levels=None, n_samples=256, batch_size=4096)
        n, n_stacks, kw = 128, 3, dict(motion_deg=3.0, motion_mm=1.5)
    torch.manual_seed(0)
    sim = dict(res_r=1.0, res_s=1.0, gap=3.0)
    sim.update(kw)
    slices, _, _ = simulate_slices(n=n, n_stacks=n_stacks, device=dev, **sim)
    dataset = Dataset(slices, args)
    model = nb.NeSVoR(dataset.transformation, dataset.resolution, dataset.mean, dataset.bounding_box, args)
    trainer = FusedTrainer(model, args)
    st = trainer.state
    B, S = args.batch_size, args.n_samples
    for _ in range(20):  # a few real iterations so that the table is not at its init
        trainer.step(**dataset.get_batch(B, dev))
    batch = dataset.get_batch(B, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    batch_random = batch
    key = (batch["slice_idx"].long() << 26) | (((batch["xyz"][:, 1] - dataset.xyz[:, 1].min()).round().long().clamp(0, 8191)) << 13) | (batch["xyz"][:, 0] - dataset.xyz[:, 0].min()).round().long().clamp(0, 8191)
    order = torch.argsort(key)
    batch_sorted = {k: v[order].contiguous() for k, v in batch.items()}
    for var in [(v, ab, loc, gr) for v in a.variants.split(",") for ab in a.ablate.split(",") for loc in a.locality.split(",") for gr in a.groups.split(",")]:
        var, ablate, loc, gr = var
        _lib.check(_lib.lib().nsv_set_fused_tc_groups(int(gr)))
        batch = batch_sorted if int(loc) else batch_random
        _lib.lib().nsv_set_fused_tile_order(int(loc))
        os.environ["NSV_ABLATE"] = ablate
        agg, fast, *rest = (int(x) for x in var.split(":"))
        smem = rest[0] if rest else -2  # third field: levels staged in shared memory (-1 as many as fit, 0 none)
        _lib.set_fused_tuning(agg, fast)
        _lib.lib().nsv_set_fused_smem_levels(smem)
        durs = []
        for i in range(3 + a.reps):
            flush.zero_()
            st.grad.zero_()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record(stream)
            st.forward_backward(batch["xyz"], batch["v"], batch["slice_idx"], None, seed=0, offset=i * B * S)
            k1.record(stream)
            torch.cuda.synchronize()
            if i >= 3:
                durs.append(k0.elapsed_time(k1))
        print(json.dumps({"cfg": a.cfg, "impl": a.impl, "agg_max": agg, "fast": fast, "smem_levels": smem, "locality": int(loc), "groups": int(gr), "ablate": int(ablate), "ms_mean": sum(durs) / len(durs), "ms_min": min(durs),
                          "gq_per_s": B * S / (min(durs) * 1e-3) / 1e9}), flush=True)
    _lib.set_fused_tuning(-1, -1)
    _lib.lib().nsv_set_fused_smem_levels(-2)
    _lib.lib().nsv_set_fused_tile_order(-1)
    _lib.lib().nsv_set_fused_tc_groups(-1)
    batch = batch_random
    if args.n_levels_bias:  # the mean(log_bias) pre-pass alone (part of every forward_backward timed above)
        import ctypes

        prm = st.params_struct()
        durs = []
        for i in range(3 + a.reps):
            flush.zero_()
            st.losses.zero_()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record(stream)
            rc = _lib.lib().nsv_inr_bias_mean(ctypes.byref(st.cfg), ctypes.byref(prm), _lib.ptr(batch["xyz"]), _lib.ptr(batch["slice_idx"]),
                                              ctypes.c_void_p(0), ctypes.c_uint64(0), ctypes.c_uint64(i * B * S),
                                              ctypes.c_void_p(st.losses.data_ptr() + 16), ctypes.c_int64(B), ctypes.c_int(S), _lib.stream(dev))
            k1.record(stream)
            torch.cuda.synchronize()
            _lib.check(rc, "nsv_inr_bias_mean")
            if i >= 3:
                durs.append(k0.elapsed_time(k1))
        print(json.dumps({"cfg": a.cfg, "kernel": "nsv_inr_bias_mean", "ms_mean": sum(durs) / len(durs), "ms_min": min(durs)}), flush=True)
    if a.timers:
        import ctypes

        os.environ["NSV_ABLATE"] = "0"
        counters = torch.zeros(16, dtype=torch.int64, device=dev)
        _lib.lib().nsv_set_fused_timers(ctypes.c_void_p(counters.data_ptr()))
        n = 10
        for i in range(n):
            flush.zero_()
            st.grad.zero_()
            st.forward_backward(batch["xyz"], batch["v"], batch["slice_idx"], None, seed=0, offset=i * B * S)
        torch.cuda.synchronize()
        _lib.lib().nsv_set_fused_timers(ctypes.c_void_p(0))
        c = [float(x) for x in counters.cpu()]
        if a.impl == "ws":
            mem = dict(zip(["gather", "wait_x_empty", "scatter", "wait_dx_full"], [x / (n * 148 * 16) for x in c[:4]]))
            chain = dict(zip(["wait_x_full", "mma_wait", "epilogues", "barriers", "losses", "wait_dx_empty"], [x / (n * 148 * 8) for x in c[8:14]]))
            print(json.dumps({"mem_warp_cycles": mem, "mem_total": sum(mem.values()), "chain_warp_cycles": chain, "chain_total": sum(chain.values())}), flush=True)
        else:
            names = ["gather", "barriers", "mma_wait", "epilogues", "losses", "scatter", "pixel_barrier", "-"]
            per_warp = [x / (n * 148 * 16) for x in c[:8]]
            print(json.dumps({"phase_cycles_per_warp": dict(zip(names, per_warp)), "total": sum(per_warp)}), flush=True)


if __name__ == "__main__":
    main()
