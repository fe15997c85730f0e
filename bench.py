#!/usr/bin/env python
"""bench.py -- INR training throughput of the NeSVoR reconstruction hot path on B200.

Metric (BASELINE.json): INR training samples/sec at batch 2^20 -- hash-grid queries (pixel x PSF
sample) fully processed forward + backward + optimiser per second.  Workload = BASELINE config 2:
3-D Shepp-Logan phantom 128^3, 3 orthogonal simulated stacks (225^2 x 77 slices, 1 mm in-plane,
3 mm thick), 16-level hash grid T=2^19 F=2, 64-wide MLP with 3 hidden layers ("4-layer"),
n-samples 128, batch 8192 pixels => 2^20 queries / iteration / GPU, density + slice-scale heads
(--no-pixel-variance --no-slice-variance --no-transformation-optimization), edge regulariser.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  Keys beyond the base contract: "roofline" (kernel A alone,
algorithmic bytes / CUDA-event time, vs MEASURED_PEAKS.json), "cpu_baseline" (the oracle port on
the host cores, bounded sample), "e2e" (same metric through the public API with host buffers).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from argparse import Namespace

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "INR training samples/sec at batch 2^20"
UNIT = "queries/s"
WORKLOAD = dict(
    workload="BASELINE config 2: phantom 128^3, 3 orthogonal stacks, hash grid L=16 T=2^19 F=2, MLP 64x3 hidden (4 linear layers), "
             "n_samples 128, batch 8192 px = 2^20 queries/iter/GPU",
    n=128, n_stacks=3, n_levels=16, log2_hashmap_size=19, width=64, depth=3, n_samples=128, batch_size=8192)


L2_NOTE = ("per-step working set 184 MB (params+grads+Adam moments) > 126 MB L2; kernel-only timing flushes L2 with a 256 MB write")


def config_block(world, scaling="weak", batch_size=None):
    """`config` of the JSON line -- identical for both arms (the driver compares them): the workload and how it is sharded."""
    c = dict(WORKLOAD, parallelism=f"dp{world}", l2=L2_NOTE, noise="in-kernel Philox (ours) / torch.randn (reference arm)", scaling=scaling)
    if batch_size is not None:
        c["batch_size"] = batch_size
    return c


def make_args(device, **kw):
    import torch

    a = dict(n_features_per_level=2, log2_hashmap_size=19, level_scale=1.3819, coarsest_resolution=16.0, finest_resolution=0.5,
             n_levels_bias=0, depth=3, width=64, n_features_z=15, n_features_slice=16, no_transformation_optimization=True,
             no_slice_scale=False, no_pixel_variance=True, no_slice_variance=True, single_precision=False,
             weight_transformation=0.1, weight_bias=100.0, image_regularization="edge", weight_image=2.0, delta=0.2,
             learning_rate=5e-3, gamma=0.33, milestones=[0.5, 0.75, 0.9], n_iter=5000, batch_size=8192, n_samples=128,
             dtype=torch.float16, device=device, n_levels=16, base_resolution=None, seed=0, fused=True, mask_threshold=1.0)
    a.update(kw)
    return Namespace(**a)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, val in zip(names, r[3:7]) if val.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_throughput(n_pixels=8192, n_samples=128, steps=2, warmup=1, threads=None, budget_s=None):
    """The reference has no CPU path; the CPU baseline is the oracle port:
    oracle/inr_oracle.py forward + autograd backward + torch AdamW, config-2 model, all host cores.

    Same configuration as the GPU arm: every step is a FULL config-2 iteration (8192 px x 128 samples = 2^20 queries,
    ~3-6 s on 16 cores) unless `budget_s` is given and (warmup + steps) full iterations would not fit in it: then the
    FIRST timed step stays full-size and the others shrink to a bounded sample of the same workload (>= 256 px).  `value` is
    total queries / total time over the timed steps; the full-size iteration is also reported on its own."""
    import torch
    from oracle import inr_oracle as io

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    n_slices = 231
    g = torch.Generator().manual_seed(0)
    ax = torch.zeros(n_slices, 6)
    ax[:, 3:] = torch.randn(n_slices, 3, generator=g) * 20
    res = torch.tensor([[1.0, 1.0, 3.0]]).repeat(n_slices, 1)
    bb = torch.tensor([[-66.0, -66.0, -66.0], [66.0, 66.0, 66.0]])
    cfg = io.INRConfig(n_levels=16, base_resolution=9, level_scale=1.3819, log2_hashmap_size=19, width=64, depth=3,
                       no_transformation_optimization=True, no_pixel_variance=True, no_slice_variance=True, n_samples=n_samples,
                       delta=0.2 * 0.3, emulate_fp16=False)
    om = io.OracleNeSVoR(cfg, n_slices, ax, res, bb)
    opt = io.make_optimizer(om)

    def step(n_px):
        xyz = (torch.rand(n_px, 3, generator=g) - 0.5) * 100
        xyz[:, 2] = 0
        v = torch.rand(n_px, generator=g)
        idx = torch.randint(0, n_slices, (n_px,), generator=g)
        t0 = time.perf_counter()
        noise = torch.randn(n_px, n_samples, 3, generator=g)
        losses = om.forward(xyz, v, idx, noise)
        om.total_loss(losses).backward()
        opt.step()
        opt.zero_grad()
        return time.perf_counter() - t0

    small = n_pixels
    step(256)  # untimed: thread pool / allocator warm-up
    n_warm = warmup
    if budget_s is not None:  # ONE full-size warm-up iteration is the yardstick: do warmup + steps of them fit in the budget?
        t_full = step(n_pixels)
        n_warm = max(warmup - 1, 0)
        if (n_warm + steps) * t_full > budget_s - t_full:
            per_step = max(budget_s - 2 * t_full, 0.25 * budget_s) / max(n_warm + steps - 1, 1)
            small = int(max(256, min(n_pixels, n_pixels * per_step / t_full * 0.5)) // 256 * 256)  # small batches run less efficiently
    for _ in range(n_warm):
        step(small)
    sizes = [n_pixels] + [small] * (steps - 1)
    times = [step(n) for n in sizes]
    total_q = sum(sizes) * n_samples
    return {"value": total_q / sum(times), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{steps} timed iteration(s) of the config-2 model, oracle/inr_oracle.py fwd+bwd+AdamW, torch {torch.__version__}, {threads} threads: "
                      f"1 x {n_pixels} px x {n_samples} samples (the full 2^20-query batch, {times[0]:.2f} s)"
                      + (f" + {steps - 1} x {small} px x {n_samples}" if steps > 1 else ""),
            "ms_per_step": sum(times) / len(times) * 1e3,
            "full_batch_iteration": {"queries": n_pixels * n_samples, "seconds": times[0], "queries_per_s": n_pixels * n_samples / times[0]},
            "all_steps_full_batch": small == n_pixels}


def run_reference_arm(a):
    """`--impl reference`: the reference's algorithm for the path on the host cores (the oracle port -- the reference itself has
    no CPU path and its tiny-cuda-nn dependency is absent), same metric, same config-2 batch: every step is a
    full 8192 px x 128 samples iteration as long as warmup + steps of them fit in ~4 minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_oracle_throughput(steps=max(a.steps, 1), warmup=max(a.warmup, 0), budget_s=240.0)
    line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "impl": "reference", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_block(a.gpus, a.scaling),
            "notes": "reference arm = CPU oracle port (oracle/inr_oracle.py); the reference ships no CPU path and its tcnn dependency is absent",
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "full_batch_iteration", "all_steps_full_batch")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ kernel B
def kernel_b_leg(dev, flush, reps=5):
    """slice_acquisition forward (kernel B) on the config-2 stacks: 231 slices x 225^2 pixels gathered from the 128^3
    phantom through the (1, 1, 3)-ratio PSF.  Algorithmic bytes per slice pixel = taps_nnz * 8 corners * 4 B read + 4 B
    written; reported against the same HBM peak as kernel A (the 8 MB volume is L2-resident)."""
    import torch
    from nesvor_b200.data.phantom import STACK_ORIENTATIONS, phantom3d, stack_axisangles, stack_geometry
    from nesvor_b200.slice_acquisition import slice_acquisition
    from nesvor_b200.transform import RigidTransform, mat_update_resolution
    from nesvor_b200.utils import get_PSF

    n = WORKLOAD["n"]
    ss, n_slice = stack_geometry(n, 1.0, 1.0, 3.0)
    vol = torch.tensor(phantom3d(n), dtype=torch.float32, device=dev)[None, None]
    psf = get_PSF(res_ratio=(1.0, 1.0, 3.0), device=dev)
    ax = stack_axisangles(STACK_ORIENTATIONS[: WORKLOAD["n_stacks"]], n_slice, 3.0).to(dev)
    mat = mat_update_resolution(RigidTransform(ax, trans_first=True).matrix(), 1, 1.0).contiguous()
    taps = int((psf != 0).sum())
    durs = []
    for i in range(2 + reps):
        flush.zero_()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        out = slice_acquisition(mat, vol, None, None, psf, (ss, ss), 1.0, False, False)
        k1.record()
        torch.cuda.synchronize()
        if i >= 2:
            durs.append(k0.elapsed_time(k1))
    ms = sum(durs) / len(durs)
    n_px = out.numel()
    alg = n_px * (taps * 8 * 4 + 4)
    return {"op": "slice_acquisition forward (fp32, linear interpolation, incl. the zero-filled output allocation)", "slices": int(out.shape[0]), "slice_shape": [ss, ss], "psf_taps": taps,
            "ms": ms, "pixels_per_s": n_px / (ms * 1e-3), "algorithmic_bytes": alg, "achieved": alg / (ms * 1e-3) / 1e9, "unit": "GB/s"}


def render_leg(dev, model, args, flush, reps=5):
    """Inference side of the path (sample.py:17-53, `sample_points` / `sample_slice`): the forward-only fused kernel
    nsv_inr_render on one inference batch of `nesvor reconstruct` for this configuration -- 8 x batch_size points x
    2 x n_samples PSF samples (cli/commands.py:94-97) = 16.8 M queries per launch.  Algorithmic bytes per query
    (SURVEY s.8d): L x 8 x F x 2 B = 512 B of fp16 table gathers."""
    import torch
    from nesvor_b200.nesvor.fused import attach_render_state, fused_render

    try:
        M, S = 8 * args.batch_size, 2 * args.n_samples
        st = attach_render_state(model.inr, args)
        bb = model.inr.bounding_box
        g = torch.Generator(device=dev).manual_seed(3)
        xyz = bb[0] + (bb[1] - bb[0]) * (0.1 + 0.8 * torch.rand(M, 3, device=dev, generator=g))
        durs = []
        for i in range(2 + reps):
            flush.zero_()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            out = fused_render(model.inr, xyz, None, 0.4247, S, state=st)
            k1.record()
            torch.cuda.synchronize()
            if i >= 2:
                durs.append(k0.elapsed_time(k1))
        ms = sorted(durs)[len(durs) // 2]
        L = st.cfg.grid.n_levels
        alg = M * S * L * 8 * 2 * 2
        return {"op": "nsv_inr_render (forward only: sample generation, hash-grid gather, MLP, softplus, PSF mean)", "points": M, "n_samples": S,
                "queries": M * S, "ms": ms, "queries_per_s": M * S / (ms * 1e-3), "algorithmic_bytes": alg, "achieved": alg / (ms * 1e-3) / 1e9,
                "unit": "GB/s", "finite": bool(torch.isfinite(out).all())}
    except Exception as e:  # informative leg
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def other_heads_leg(dev, steps=50):
    """Whole fused iteration (kernel A [+ the mean(log_bias) pre-pass] + finalize + transReg + AdamW, device-resident batches)
    at 2^20 queries for the head configurations of BASELINE configs 3 and 5 -- informative only; the bench line's
    `value` stays config 2.  cfg3: 128^3 phantom, 3 stacks with injected motion, reference-default heads (sigma_net,
    slice scale / variance, pose optimisation), S = 256.  cfg5: 138^3 phantom at 0.8 mm, 9 stacks x 30 slices, the same
    heads + bias field on 4 levels (b_net), finest resolution 0.5."""
    import torch
    import nesvor_b200 as nb
    from nesvor_b200.data.phantom import simulate_slices
    from nesvor_b200.nesvor.fused import FusedTrainer, HostBatchFeeder
    from nesvor_b200.nesvor.train import Dataset

    heads = dict(depth=1, no_pixel_variance=False, no_slice_variance=False, no_transformation_optimization=False, n_levels=None,
                 n_samples=256, batch_size=4096)
    cases = {"cfg3_heads": (dict(heads), dict(n=128, n_stacks=3, res_r=1.0, res_s=1.0, gap=3.0, motion_deg=3.0, motion_mm=1.5)),
             "cfg5_heads": (dict(heads, n_levels_bias=4, finest_resolution=0.5), dict(n=138, n_stacks=9, res_r=0.8, res_s=0.8, gap=3.0, n_slice=30))}
    out = {}
    for name, (kw, sim) in cases.items():
        try:
            args = make_args(dev, **kw)
            torch.manual_seed(0)
            slices, _, _ = simulate_slices(device=dev, **sim)
            dataset = Dataset(slices, args)
            model = nb.NeSVoR(dataset.transformation, dataset.resolution, dataset.mean, dataset.bounding_box, args)
            trainer = FusedTrainer(model, args)
            B, S = args.batch_size, args.n_samples
            for _ in range(10):
                trainer.step(**dataset.get_batch(B, dev))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                losses = trainer.step(**dataset.get_batch(B, dev))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"ms_per_step": ms, "queries_per_s": B * S / (ms * 1e-3), "n_levels": int(trainer.state.cfg.grid.n_levels),
                         "n_slices": int(model.n_slices), "batch_size": B, "n_samples": S,
                         "losses_last_step": {k: float(v) for k, v in losses.items()}}
            del trainer, model, dataset, slices
        except Exception as e:  # informative leg: never take the bench line down with it
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
    return out


def unfused_leg(dev, dataset, steps=5):
    """The same config-2 iteration on the UNFUSED native path -- `NeSVoR.forward` composed from the standalone CUDA ops
    (hash grid, fused MLP, pose converters) under autograd + torch.optim.AdamW + GradScaler, i.e. the reference's own loop
    structure (train.py:179-198) with one kernel per op and activations round-tripping HBM.  Informative only: it is what
    the fused kernel A is measured against on the same GPU (SURVEY s.8d "unfused GPU baseline")."""
    import torch
    import nesvor_b200 as nb
    from nesvor_b200.nesvor.train import build_optimizer, loss_weights

    try:
        args = make_args(dev, fused=False)
        torch.manual_seed(0)
        model = nb.NeSVoR(dataset.transformation, dataset.resolution, dataset.mean, dataset.bounding_box, args)
        opt = build_optimizer(model, args)
        scaler = torch.amp.GradScaler("cuda", init_scale=1.0, enabled=True, growth_factor=2.0, backoff_factor=0.5)
        wts = loss_weights(args)
        B, S = args.batch_size, args.n_samples

        def it():
            losses = model(**dataset.get_batch(B, dev))
            loss = sum(wts[k] * v for k, v in losses.items() if k in wts and wts[k])
            scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()
            opt.zero_grad()

        for _ in range(2):
            it()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            it()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"ms_per_step": ms, "queries_per_s": B * S / (ms * 1e-3), "steps": steps,
                "what": "NeSVoR.forward from standalone native ops under autograd + torch.optim.AdamW (config 2, 2^20 queries)"}
    except Exception as e:  # informative leg: never take the bench line down with it
        return {"error": f"{type(e).__name__}: {e}"[:300]}


# ------------------------------------------------------------------------------------------ our arm
def run_ours(a):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from nesvor_b200.csrc import build as nsv_build

    nsv_build.build()
    import nesvor_b200 as nb
    from nesvor_b200 import _lib as nsv_lib

    nsv_lib.set_fused_impl(a.fused_impl)
    from nesvor_b200.data.phantom import simulate_slices
    from nesvor_b200.nesvor.fused import FusedTrainer, HostBatchFeeder
    from nesvor_b200.nesvor.train import Dataset

    strong = a.scaling == "strong"
    if strong:  # BASELINE config 4: 2^22 queries / iteration GLOBALLY (32768 px x 128), sharded over the ranks
        if 32768 % world:
            raise SystemExit("bench.py --scaling strong: the 32768-pixel global batch must divide by the number of ranks")
        args = make_args(dev, batch_size=32768 // world)
    else:
        args = make_args(dev)
    torch.manual_seed(0)
    slices, _, _ = simulate_slices(n=WORKLOAD["n"], n_stacks=WORKLOAD["n_stacks"], res_r=1.0, res_s=1.0, gap=3.0, device=dev)
    dataset = Dataset(slices, args)
    model = nb.NeSVoR(dataset.transformation, dataset.resolution, dataset.mean, dataset.bounding_box, args)
    trainer = FusedTrainer(model, args)
    st = trainer.state
    B, S = args.batch_size, args.n_samples
    n_q = B * S
    # every rank draws from its own shuffled copy of the pixel table (weak scaling: B pixels per rank)
    torch.manual_seed(1234 + rank)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one_step(batch):
        if world == 1:
            return trainer.step(**batch)
        return trainer.step_distributed(dist, world, **batch)

    # ---------------- value: device-resident inputs, whole iteration (fwd + bwd + AdamW) ----------
    n_warm = max(a.warmup, 20)  # at least 20 untimed iterations (16 ms): clocks and the L2-resident table settle before the timed region
    for _ in range(n_warm):
        one_step(dataset.get_batch(B, dev))
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local)  # sampled over every timed leg below (value, e2e, kernel-only): the value leg alone is < 1 s
    clocks.__enter__()
    torch.cuda.nvtx.range_push("nsv_timed")  # ncu --nvtx --nvtx-include "nsv_timed/" lists exactly these launches
    ev0.record()
    for _ in range(a.steps):
        losses = one_step(dataset.get_batch(B, dev))
    ev1.record()
    torch.cuda.nvtx.range_pop()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * n_q * a.steps / (ms_total * 1e-3)

    # ---------------- e2e: host buffers through the public step API ----------------
    host = []
    for _ in range(4):
        b = dataset.get_batch(B, dev)
        host.append({k: v.cpu().pin_memory() for k, v in b.items()})
    # H2D of EVERY step's inputs from pinned memory, inside the timed region: nesvor_b200's HostBatchFeeder copies batch i + 1 on a
    # side stream while the kernels of batch i run (three copies per step: xyz, v, slice_idx)
    feeder = HostBatchFeeder(dev)

    def e2e_steps(n):
        pending, got = None, {}
        for batch in feeder.feed(host[i % len(host)] for i in range(n)):
            out = one_step(batch)
            # D2H of the step's losses, every step, the way nesvor_b200.train() does it: ONE async 32-byte copy into pinned
            # memory behind the step's kernels, read after the NEXT step has been enqueued (LossHandle) -- the reference's loop
            # drains the GPU with 5-6 `.item()` calls per iteration (train.py:199-200)
            handle = trainer.losses_to_host(out)
            if pending is not None:
                got = pending.get()
            pending = handle
        return pending.get(), pending.nbytes

    e2e_steps(n_warm)  # untimed warm-up of THIS path (side streams, pinned buffers and feeder slots are set up here)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h2d0 = feeder.bytes_copied
    e0.record()
    loss_host, d2h = e2e_steps(a.steps)
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n_q * a.steps / (float(t.item()) * 1e-3)
    h2d = (feeder.bytes_copied - h2d0) // a.steps  # counted from the tensors copied: B x (xyz 12 + v 4 + slice_idx 8) bytes
    assert h2d == B * (3 * 4 + 4 + 8), h2d

    # ---------------- exchange step alone (N > 1): gradient mean + AdamW + parameter refresh over the ranks ----------------
    exchange_ms = None
    if world > 1:
        sync_all()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_x = 20
        x0.record()
        for _ in range(n_x):
            trainer.iteration += 1
            trainer._dp_update(dist, world)  # zero gradients: the parameters only see the weight decay of 20 tiny steps
        x1.record()
        sync_all()
        t = torch.tensor([x0.elapsed_time(x1) / n_x], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        exchange_ms = float(t.item())

    if rank != 0:
        clocks.__exit__(None, None, None)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline: kernel A alone, L2 flushed between launches ----------------
    L = st.cfg.grid.n_levels
    alg_bytes = n_q * L * 8 * 2 * (2 + 4)  # fp16 table gather + fp32 gradient scatter, per SURVEY s.8d
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    batch = dataset.get_batch(B, dev)
    durs = []
    stream = torch.cuda.current_stream()
    for i in range(3 + min(a.steps, 20)):
        flush.zero_()
        st.grad.zero_()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        st.forward_backward(batch["xyz"], batch["v"], batch["slice_idx"], None, seed=0, offset=i * n_q)
        k1.record(stream)
        torch.cuda.synchronize()
        if i >= 3:
            durs.append(k0.elapsed_time(k1))
    k_ms = sum(durs) / len(durs)
    kernel_b = kernel_b_leg(dev, flush)
    render = render_leg(dev, model, args, flush) if world == 1 else None
    other_heads = other_heads_leg(dev) if world == 1 else None
    if world == 1:  # kernel B next to the reference's own CUDA extension on this GPU (subprocess: foreign kernels stay out of this context)
        try:
            import subprocess

            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "kernel_b_vs_reference.py"), "--reps", "5"],
                               capture_output=True, text=True, timeout=300)
            ref_line = next((l for l in reversed(r.stdout.splitlines()) if l.startswith("{")), None)
            kernel_b["vs_reference_cuda_extension"] = json.loads(ref_line) if ref_line else {"available": False, "why": (r.stderr or "no output")[-200:]}
        except Exception as e:
            kernel_b["vs_reference_cuda_extension"] = {"available": False, "why": f"{type(e).__name__}: {e}"[:200]}
    clocks.__exit__(None, None, None)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    kname = {"mma": "inr_train_kernel<64,3,false,128> (mma.sync)", "ws": "inr_train_ws_kernel<3,false,false> (tcgen05/TMEM, warp-specialised)"}.get(
        a.fused_impl, "inr_train_tc_kernel<3,false> (tcgen05/TMEM)")
    roofline = {"bound": "hbm", "kernel": kname + " + 1-block finalize", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "kernel_ms": k_ms, "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                "kernel_queries_per_s": n_q / (k_ms * 1e-3)}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get("inr_train_kernel_dram_bytes_per_launch")
        except Exception:
            pass

    cpu = cpu_oracle_throughput(steps=2, warmup=1) if world == 1 else None  # rank 0 at N = 1 only: two full config-2 iterations
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "untimed_iterations_before_each_timed_leg": n_warm,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f16 operands / f32 accumulate",
            "data": "synthetic", "config": config_block(world, a.scaling, B if strong else None),
            "workload_detail": {"n_pixels_in_table": int(dataset.xyz.shape[0]), "n_slices": model.n_slices, "queries_per_rank_per_step": n_q,
                                "global_queries_per_step": n_q * world, "smem_staged_levels": "levels gathered from the TMA-staged shared-memory copy: see DESIGN.md s.4"},
            "dp": {"mode": trainer.dp_mode, "multimem": bool(getattr(trainer, "dp_multimem_active", False)),
                   "autotune_ms": getattr(trainer, "dp_autotune_ms", None), "exchange_ms": exchange_ms,
                   "what": "exchange = gradient mean over the ranks + AdamW + fp16 parameter refresh; 'peer' = one fused kernel over NVLink peer memory "
                           "(nsv_adamw_step_dp; multimem: sum and replication inside the NVSwitch through multicast addresses, nsv_adamw_step_dp_mc), "
                           "'allreduce' = NCCL all-reduce + nsv_adamw_step"} if world > 1 else None,
            "clocks": clocks.summary(), "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": 3 * a.steps, "roofline": roofline, "kernel_b": dict(kernel_b, frac=kernel_b["achieved"] / peak, peak=peak),
            "losses_last_step": {k: float(v) for k, v in losses.items()}}
    if render is not None:
        line["render"] = dict(render, frac=render["achieved"] / peak) if "achieved" in render else render
    if cpu is not None:
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "full_batch_iteration")}
    if other_heads is not None:
        line["other_heads"] = other_heads
        line["unfused_gpu"] = unfused_leg(dev, dataset)  # last: a failure here cannot touch the numbers above
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 2^20 queries per rank per step (default); strong: BASELINE config 4, 2^22 queries per step globally")
    ap.add_argument("--fused-impl", default="auto", choices=["auto", "mma", "tcgen05", "ws"], help="implementation of kernel A")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
