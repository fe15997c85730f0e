#!/usr/bin/env python
"""Diagnostic: trains the config-3p workload (poses fixed or joint) with FusedTrainer and reports the first iteration at which
a loss or a parameter segment stops being finite, with the loss trajectory before it.  python tools/diag_nan.py [--fixed 1]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fixed", type=int, default=1)
    ap.add_argument("--iters", type=int, default=3000)
    ap.add_argument("--impl", default="auto")
    ap.add_argument("--sequence", default="", help="train mode: comma list of `fixed` values run one after the other in THIS process")
    ap.add_argument("--mode", default="step", choices=["step", "train"], help="train: through nesvor_b200.train + the evaluation of psnr_phantom.run_pose_recovery")
    a = ap.parse_args()
    import torch

    import nesvor_b200 as nb
    import psnr_phantom as pp
    from nesvor_b200 import _lib
    from nesvor_b200.data.phantom import simulate_slices
    from nesvor_b200.nesvor.fused import FusedTrainer
    from nesvor_b200.nesvor.train import Dataset

    dev = torch.device("cuda", 0)
    _lib.set_fused_impl(a.impl)
    c = pp.CFGS["3p"]
    args = pp.make_args(dev, n_iter=a.iters, batch_size=4096, n_samples=64, **dict(c["args"], no_transformation_optimization=bool(a.fixed)))
    torch.manual_seed(0)
    slices, _, _ = simulate_slices(device=dev, **c["sim"])
    if a.mode == "train" and a.sequence:
        reps = []
        for f in a.sequence.split(","):
            args = pp.make_args(dev, n_iter=a.iters, batch_size=4096, n_samples=64, no_loss_sync=True, **dict(c["args"], no_transformation_optimization=bool(int(f))))
            torch.manual_seed(0)
            slices, volume, _ = simulate_slices(device=dev, **c["sim"])
            from nesvor_b200.nesvor.fused import attach_render_state, fused_render

            inr, out_slices, mask = nb.train(slices, args)
            grid = pp.phantom_grid(c["sim"]["n"], c["sim"]["res_r"])
            st = attach_render_state(inr, args)
            rec = torch.cat([fused_render(inr, grid[i : i + (1 << 18)].to(dev), None, 0.0, 1).cpu() for i in range(0, grid.shape[0], 1 << 18)])
            gt = volume[0, 0].reshape(-1).cpu()
            reps.append({"fixed": int(f), "params_finite": {k: bool(torch.isfinite(v).all()) for k, v in inr.state_dict().items()},
                         "render_nonfinite": int((~torch.isfinite(rec)).sum()), "gt_nonfinite": int((~torch.isfinite(gt)).sum()),
                         "psnr_inside": pp.psnr(rec, gt, gt > 0), "n_inside": int((gt > 0).sum()), "rec_absmax": float(rec[torch.isfinite(rec)].abs().max())})
        print(json.dumps({"mode": "train-sequence", "runs": reps}))
        return
    if a.mode == "train":
        from nesvor_b200.nesvor.fused import attach_render_state, fused_render

        args.no_loss_sync = True
        inr, out_slices, mask = nb.train(slices, args)
        rep = {"mode": "train", "fixed": a.fixed, "params_finite": {k: bool(torch.isfinite(v).all()) for k, v in inr.state_dict().items()},
               "params_absmax": {k: float(v.float().abs().max()) for k, v in inr.state_dict().items()}}
        grid = pp.phantom_grid(c["sim"]["n"], c["sim"]["res_r"])
        st = attach_render_state(inr, args)
        rep["flat16_finite"] = bool(torch.isfinite(st.flat16).all())
        rec = torch.cat([fused_render(inr, grid[i : i + (1 << 18)].to(dev), None, 0.0, 1).cpu() for i in range(0, grid.shape[0], 1 << 18)])
        bad = ~torch.isfinite(rec)
        rep["render_nonfinite"] = int(bad.sum())
        rep["render_absmax_finite"] = float(rec[~bad].abs().max())
        if bad.any():
            idx = bad.nonzero().flatten()[:5]
            rep["first_bad_points"] = grid[idx].tolist()
            rep["bbox"] = inr.bounding_box.tolist()
            x = grid[idx].to(dev)
            rep["unfused_forward_at_bad_points"] = inr(x[:, None], False).flatten().tolist()
        print(json.dumps(rep))
        return
    ds = Dataset(slices, args)
    model = nb.NeSVoR(ds.transformation, ds.resolution, ds.mean, ds.bounding_box, args)
    tr = FusedTrainer(model, args)
    st = tr.state
    hist = []
    milestones = [int(m * a.iters) for m in args.milestones]
    for it in range(1, a.iters + 1):
        losses = tr.step(**ds.get_batch(args.batch_size, dev))
        if it in milestones:
            tr.decay_lr(args.gamma)
        if it % 25 == 0 or it < 5:
            vals = {k: float(v) for k, v in losses.items()}
            segs = {n: bool(torch.isfinite(st.seg(n)).all()) for n in st.offsets}
            extra = {"lvs_minmax": [float(st.seg("log_var_slice").min()), float(st.seg("log_var_slice").max())] if "log_var_slice" in st.offsets else None,
                     "table_absmax": float(st.seg("table").abs().max()), "mlp_absmax": float(st.seg("mlp").abs().max()),
                     "m_absmax": float(tr.exp_avg.abs().max()), "v_max": float(tr.exp_avg_sq.max())}
            hist.append({"it": it, "losses": vals, "finite": segs, **extra})
            if not all(segs.values()) or not all(v == v and abs(v) != float("inf") for v in vals.values()):
                break
    print(json.dumps({"fixed": a.fixed, "impl": a.impl, "stopped_at": hist[-1]["it"], "tail": hist[-6:]}))


if __name__ == "__main__":
    main()
