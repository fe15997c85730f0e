"""Where the end-to-end iteration (host batches in, loss values out) spends its time beyond the device-resident one:
host enqueue time per iteration (perf_counter around the loop body, no synchronisation) and device time per iteration for the
combinations {in-stream H2D copies, HostBatchFeeder} x {losses read 1 / 2 iterations late}.  Config 2, one GPU.
    python tools/e2e_pipeline_check.py [--steps 200] > profiles/...json"""
import argparse
import cProfile
import io
import json
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    a = ap.parse_args()
    import bench
    import nesvor_b200 as nb
    from nesvor_b200.data.phantom import simulate_slices
    from nesvor_b200.nesvor.fused import FusedTrainer, HostBatchFeeder
    from nesvor_b200.nesvor.train import Dataset

    dev = torch.device("cuda", 0)
    args = bench.make_args(dev)
    torch.manual_seed(0)
    slices, _, _ = simulate_slices(n=bench.WORKLOAD["n"], n_stacks=bench.WORKLOAD["n_stacks"], res_r=1.0, res_s=1.0, gap=3.0, device=dev)
    dataset = Dataset(slices, args)
    model = nb.NeSVoR(dataset.transformation, dataset.resolution, dataset.mean, dataset.bounding_box, args)
    trainer = FusedTrainer(model, args)
    B = args.batch_size
    dev_batches = [dataset.get_batch(B, dev) for _ in range(4)]
    dev_batches = [{k: v.clone() for k, v in b.items()} for b in dev_batches]
    host = [{k: v.cpu().pin_memory() for k, v in b.items()} for b in dev_batches]
    for b in dev_batches:
        trainer.step(**b)
    torch.cuda.synchronize()
    out = {"steps": a.steps, "what": __doc__.split("\n")[0]}

    def timed(name, body):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        host_busy = body()
        e1.record()
        torch.cuda.synchronize()
        out[name] = {"device_ms_per_step": e0.elapsed_time(e1) / a.steps, "wall_ms_per_step": (time.perf_counter() - t0) * 1e3 / a.steps,
                     "host_enqueue_ms_per_step": None if host_busy is None else host_busy * 1e3 / a.steps}

    def resident():
        t = 0.0
        for i in range(a.steps):
            t0 = time.perf_counter()
            trainer.step(**dev_batches[i % 4])
            t += time.perf_counter() - t0
        return t

    feeder = HostBatchFeeder(dev)  # one feeder for every run: stream, slots and events are set up once (warm-up below)

    def e2e(use_feeder, late, steps=None):
        def body():
            pend, busy = [], 0.0
            src = (host[i % 4] for i in range(steps or a.steps))
            it = feeder.feed(src) if use_feeder else ({k: v.to(dev, non_blocking=True) for k, v in hb.items()} for hb in src)
            t0 = time.perf_counter()
            for batch in it:
                losses = trainer.step(**batch)
                pend.append(trainer.losses_to_host(losses))
                busy += time.perf_counter() - t0
                if len(pend) > late:
                    pend.pop(0).get()
                t0 = time.perf_counter()
            for h in pend:
                h.get()
            return busy
        return body

    timed("device_resident_no_readback", resident)
    for rep in (1, 2):  # twice: one-time set-up costs (pinned buffers, side stream, allocator pools) would show as a difference
        for use_feeder in (False, True):
            for late in (1, 2):
                e2e(use_feeder, late, steps=20)()  # untimed warm-up of exactly this path
                timed(f"{'feeder' if use_feeder else 'instream'}_read{late}_late_run{rep}", e2e(use_feeder, late))
    # host profile of the enqueue path (feeder, 2 late)
    pr = cProfile.Profile()
    pr.enable()
    e2e(True, 2)()
    pr.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    out["cprofile_feeder_read2_late"] = s.getvalue().splitlines()[:60]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
