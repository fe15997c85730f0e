"""End-to-end GPU test of the path's callers: simulated stacks -> Dataset -> train()
with the fused kernel -> Dataset.mask -> sample_volume / sample_slices on the forward-only fused renderer.
Reference flow: nesvor/nesvor/train.py:123-232, sample.py:10-64."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))


@pytest.fixture(scope="module")
def trained(native_lib):
    import nesvor_b200 as nb
    import psnr_phantom as pp
    from nesvor_b200.data.phantom import simulate_slices

    dev = torch.device("cuda", 0)
    # the synthetic stacks are lattice-aligned (pixel centres on half-integers): round() piles them into every other
    # voxel, which lifts the reference's count-based mask threshold above the blurred counts -> use a lower threshold
    args = pp.make_args(dev, n_iter=2000, batch_size=2048, n_samples=64, mask_threshold=0.1, no_loss_sync=True,
                        output_resolution=1.0, inference_batch_size=1 << 14, n_inference_samples=128, no_output_psf=False)
    torch.manual_seed(0)
    slices, volume, _ = simulate_slices(device=dev, n=48, n_stacks=3, res_r=1.0, res_s=1.0, gap=2.0)
    inr, out_slices, mask = nb.train(slices, args)
    return dict(args=args, slices=slices, volume=volume, inr=inr, out_slices=out_slices, mask=mask)


def test_train_returns_model_slices_and_mask(trained):
    t = trained
    assert len(t["out_slices"]) == len(t["slices"])
    m = t["mask"]
    frac = float(m.mask.float().mean())
    assert 0.02 < frac < 0.9, frac  # the head, not nothing and not the whole padded box
    # the mask covers the phantom's support
    import psnr_phantom as pp

    grid = pp.phantom_grid(48, 1.0).to(m.image.device)
    inside = t["volume"][0, 0].reshape(-1) > 0
    cover = (m.sample_points(grid)[inside] > 0).float().mean()
    assert cover > 0.95, float(cover)


def test_sample_volume_reconstructs_phantom(trained):
    import psnr_phantom as pp
    from nesvor_b200.nesvor.sample import sample_points, sample_volume

    t = trained
    vol = sample_volume(t["inr"], t["mask"], t["args"])
    assert torch.isfinite(vol.image).all() and int(vol.mask.sum()) > 0 and float(vol.image.max()) > 0.1
    grid = pp.phantom_grid(48, 1.0).to(vol.image.device)
    gt = t["volume"][0, 0].reshape(-1)
    import copy

    a = copy.copy(t["args"])
    a.no_output_psf = True
    rec = sample_points(t["inr"], grid, a)
    p_in = pp.psnr(rec.cpu(), gt.cpu(), (gt > 0).cpu())
    print("PSNR inside phantom after 2000 iterations:", p_in)
    assert p_in > 14.0


def test_sample_slices_fit_the_data(trained):
    from nesvor_b200.nesvor.sample import sample_slices

    t = trained
    sel = t["out_slices"][:: max(len(t["out_slices"]) // 6, 1)]
    sim = sample_slices(t["inr"], sel, t["mask"], t["args"])
    num = den = 0.0
    for s, r in zip(sel, sim):
        m = s.mask & r.mask
        assert r.image.shape == s.image.shape
        num += float(((r.image[m] - s.image[m]) ** 2).sum())
        den += float((s.image[m] ** 2).sum())
    rel = (num / max(den, 1e-30)) ** 0.5
    print("relative L2 of re-simulated slices vs input slices:", rel)
    assert den > 0 and rel < 0.35


def test_fused_and_unfused_renderers_agree(trained):
    """sample_points through nsv_inr_render (in-kernel Philox noise) vs INR.sample_batch + INR.forward (torch.randn):
    same estimator, different noise -> agreement at Monte-Carlo accuracy; without PSF they agree at fp16 accuracy."""
    import copy

    from nesvor_b200.nesvor.sample import sample_points

    t = trained
    dev = t["mask"].image.device
    xyz = (torch.rand(4096, 3, device=dev) - 0.5) * 30.0
    a_f, a_u = copy.copy(t["args"]), copy.copy(t["args"])
    a_u.fused = False
    a_f.n_inference_samples = a_u.n_inference_samples = 512  # Monte-Carlo error ~ 1 / sqrt(S)
    for no_psf, tol in ((True, 2e-3), (False, 6e-2)):
        a_f.no_output_psf = a_u.no_output_psf = no_psf
        vf, vu = sample_points(t["inr"], xyz, a_f), sample_points(t["inr"], xyz, a_u)
        rel = float((vf - vu).norm() / vu.norm())
        print("fused vs unfused renderer, no_output_psf =", no_psf, "rel-L2 =", rel)
        assert rel < tol


def test_train_with_fused_bias_field_head(native_lib):
    """BASELINE config-5 heads end to end: stacks multiplied by a smooth synthetic bias field, train() with
    n_levels_bias = 4 on the fused kernel (b_net + biasReg through nsv_inr_bias_mean).  The data term must fall,
    biasReg must stay small (its weight is 100), and the trained b_net must flow back into the nn.Module, where the
    unfused native path evaluates the same small biasReg."""
    import nesvor_b200 as nb
    import psnr_phantom as pp
    from nesvor_b200.data.phantom import simulate_slices
    from nesvor_b200.nesvor.fused import FusedTrainer
    from nesvor_b200.nesvor.train import Dataset
    from nesvor_b200.nesvor.models import NeSVoR

    dev = torch.device("cuda", 0)
    args = pp.make_args(dev, n_iter=600, batch_size=2048, n_samples=64, n_levels_bias=4, no_loss_sync=True)
    torch.manual_seed(0)
    slices, volume, _ = simulate_slices(device=dev, n=48, n_stacks=3, res_r=1.0, res_s=1.0, gap=2.0)
    for s in slices:  # smooth multiplicative field in slice coordinates, different per stack orientation
        h, w = s.image.shape[-2:]
        yy, xx = torch.meshgrid(torch.linspace(-1, 1, h, device=dev), torch.linspace(-1, 1, w, device=dev), indexing="ij")
        s.image = s.image * torch.exp(0.3 * torch.cos(1.5 * xx + 0.5 * s.stack_idx) * torch.cos(1.1 * yy))
    dataset = Dataset(slices, args)
    model = NeSVoR(dataset.transformation, dataset.resolution, dataset.mean, dataset.bounding_box, args)
    assert hasattr(model, "b_net")
    trainer = FusedTrainer(model, args)
    first, last = [], []
    for i in range(args.n_iter):
        losses = trainer.step(**dataset.get_batch(args.batch_size, dev))
        assert "biasReg" in losses
        if i < 20:
            first.append({k: float(v) for k, v in losses.items()})
        if i >= args.n_iter - 20:
            last.append({k: float(v) for k, v in losses.items()})
    mean = lambda rows, k: sum(r[k] for r in rows) / len(rows)  # noqa: E731
    print("MSE+logVar first/last:", mean(first, "MSE+logVar"), mean(last, "MSE+logVar"), "biasReg last:", mean(last, "biasReg"))
    assert all(torch.isfinite(torch.tensor(list(r.values()))).all() for r in first + last)
    assert mean(last, "MSE+logVar") < mean(first, "MSE+logVar") - 0.5
    assert mean(last, "biasReg") < 1e-2
    # parameters flow back into the nn.Module (b_net included) and the unfused native path agrees on the losses
    before = model.b_net.params.detach().clone()
    trainer.sync_to_model()
    assert not torch.equal(before, model.b_net.params.detach())
    ref = model(**dataset.get_batch(512, dev))
    assert torch.isfinite(ref["biasReg"]) and float(ref["biasReg"]) < 5e-2


def test_train_falls_back_to_the_per_op_path_when_the_fused_kernels_do_not_cover_the_configuration(native_lib, caplog):
    """`--depth 2` with the variance heads on has no fused instantiation (sigma_net with two hidden layers): `train()` with
    args.fused must warn and run the reference's loop structure (fp16 autocast forward, GradScaler, torch AdamW, train.py:179-198)
    on the per-op native kernels instead of raising; the model must still learn."""
    import logging

    import nesvor_b200 as nb
    import psnr_phantom as pp
    from nesvor_b200.data.phantom import simulate_slices

    dev = torch.device("cuda", 0)
    args = pp.make_args(dev, n_iter=60, batch_size=512, n_samples=32, depth=2, mask_threshold=0.1)
    torch.manual_seed(0)
    slices, _, _ = simulate_slices(device=dev, n=32, n_stacks=3, res_r=1.0, res_s=1.0, gap=2.0)
    with caplog.at_level(logging.WARNING):
        inr, out_slices, mask = nb.train(slices, args)
    assert any("per-op native path" in r.getMessage() for r in caplog.records)
    assert len(out_slices) == len(slices) and all(torch.isfinite(p).all() for p in inr.parameters())
    x = torch.rand(256, 3, device=dev) * 10 - 5
    assert torch.isfinite(inr(x[:, None], False)).all()


def test_host_batch_feeder_and_loss_ring(native_lib):
    """Host-resident batches through HostBatchFeeder (copies one iteration ahead on a side stream) give the same iterations as
    device-resident batches, bit for bit on the first step and within atomics noise later; the loss values come back through ONE
    32-byte copy per step (LossHandle) and equal the device values; "MSE+logVar" is the sum the finalize kernel wrote; the loss
    ring survives a wrap-around."""
    import nesvor_b200 as nb
    import psnr_phantom as pp
    from nesvor_b200.data.phantom import simulate_slices
    from nesvor_b200.nesvor import fused as F
    from nesvor_b200.nesvor.train import Dataset

    dev = torch.device("cuda", 0)
    args = pp.make_args(dev, n_iter=10, batch_size=1024, n_samples=64)
    args.no_slice_variance = False  # MSE, logVar and their sum
    torch.manual_seed(0)
    slices, _, _ = simulate_slices(device=dev, n=32, n_stacks=3, res_r=1.0, res_s=1.0, gap=2.0)
    dataset = Dataset(slices, args)
    batches = [dataset.get_batch(args.batch_size, dev) for _ in range(12)]
    batches = [{k: v.clone() for k, v in b.items()} for b in batches]
    host = [{k: v.cpu().pin_memory() for k, v in b.items()} for b in batches]

    def run(feed):
        torch.manual_seed(1)
        model = nb.NeSVoR(dataset.transformation, dataset.resolution, dataset.mean, dataset.bounding_box, args)
        tr = F.FusedTrainer(model, args)
        got, pending = [], None

        def check(h, out):
            vals = h.get()
            assert h.nbytes == 32 and set(vals) == set(out)
            for k in out:
                assert vals[k] == float(out[k]), k  # the host copy IS the device value (slots are not reused within the ring)
            assert abs(vals["MSE+logVar"] - (vals["MSE"] + vals["logVar"])) <= 1e-6 * max(1.0, abs(vals["MSE"]))
            got.append(vals)

        for batch in feed:
            out = tr.step(**batch)
            h = tr.losses_to_host(out)
            if pending is not None:
                check(*pending)  # one iteration late, like train(): the pinned ring holds 4 iterations
            pending = (h, out)
        check(*pending)
        return got, tr

    a, _ = run(batches)
    feeder = F.HostBatchFeeder(dev)
    b, tr = run(feeder.feed(host))
    assert feeder.bytes_copied == 12 * args.batch_size * (12 + 4 + 8)
    assert a[0]["MSE"] == pytest.approx(b[0]["MSE"], rel=1e-6)
    for x, y in zip(a, b):
        for k in x:
            assert y[k] == pytest.approx(x[k], rel=2e-3, abs=1e-6), k  # same batches, same seeds: atomics-order noise only
    # delivery order under overlap: every slot is overwritten only after the step that read it; tag batches by their first value
    tags = []
    slow = torch.randn(2048, 2048, device=dev)
    for batch in F.HostBatchFeeder(dev, depth=2).feed(host):
        (slow @ slow).sum()  # keeps the compute stream busy while the next copy is in flight
        tags.append(batch["v"][:4].clone())
    torch.cuda.synchronize()
    for t, hb in zip(tags, host):
        assert torch.equal(t.cpu(), hb["v"][:4])
    # wrap-around of the loss ring: values of the step after the wrap are right, the ring was cleared once
    st = tr.state
    st.loss_slot = F.LOSS_RING - 2
    o1 = tr.step(**batches[0])
    h1 = tr.losses_to_host(o1)  # its copy runs on the read-back stream: the wrap below must wait for it before clearing the ring
    o2 = tr.step(**batches[1])  # wraps: ring cleared, slot 0
    v1 = h1.get()["MSE"]
    assert st.loss_slot == 0 and float(o2["MSE"]) > 0 and float(o1["MSE"]) == 0.0 and v1 > 0
    assert float(st.loss_ring[1:].abs().sum()) == 0.0
