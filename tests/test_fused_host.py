"""CPU tests of the fused path's host side (nesvor_b200/nesvor/fused.py): how a NeSVoR model is packed into the flat
buffers kernel A reads -- segment order (trainable prefix first), 16-byte alignment, the per-head offsets that
nsv_inr_mlp_layout reports, sigma_net's re-slotted first layer, b_net (bias-field head, models.py:247-258) -- and that
parameters survive the round trip flat -> nn.Module.  No compute entry point is called (there is no GPU here)."""
from argparse import Namespace

import pytest
import torch


def make_args(**kw):
    a = dict(n_features_per_level=2, log2_hashmap_size=19, level_scale=1.3819, coarsest_resolution=16.0, finest_resolution=0.5,
             n_levels_bias=0, depth=1, width=64, n_features_z=15, n_features_slice=16, no_transformation_optimization=False,
             no_slice_scale=False, no_pixel_variance=False, no_slice_variance=False, single_precision=False,
             weight_transformation=0.1, weight_bias=100.0, image_regularization="edge", weight_image=2.0, delta=0.2,
             learning_rate=5e-3, gamma=0.33, milestones=[0.5, 0.75, 0.9], n_iter=10, batch_size=64, n_samples=128,
             dtype=torch.float16, device=torch.device("cpu"), n_levels=None, base_resolution=None, seed=0)
    a.update(kw)
    return Namespace(**a)


def build_model(args, n_slices=7):
    import nesvor_b200 as nb

    g = torch.Generator().manual_seed(3)
    ax = torch.randn(n_slices, 6, generator=g) * 0.1
    res = torch.tensor([[1.0, 1.0, 3.0]]).repeat(n_slices, 1)
    bb = torch.tensor([[-30.0, -30.0, -30.0], [30.0, 30.0, 30.0]])
    return nb.NeSVoR(nb.RigidTransform(ax, True), res, 0.7, bb, args)


@pytest.mark.parametrize("kw", [dict(), dict(n_levels_bias=4), dict(n_levels_bias=2, no_slice_variance=True),
                                dict(depth=3, no_pixel_variance=True, no_slice_variance=True, no_transformation_optimization=True)])
def test_flat_layout_and_round_trip(native_lib, kw):
    from nesvor_b200.nesvor.fused import FusedState

    args = make_args(**kw)
    model = build_model(args)
    st = FusedState(model.inr, args, model)
    W = args.width
    per_density = W * 32 + (args.depth - 1) * W * W + 16 * W
    per_head = W * 32 + (args.depth - 1) * W * W + 16 * W
    assert st.off_density == 0 and st.off_sigma == per_density
    assert st.off_bias == per_density + (0 if args.no_pixel_variance else per_head)
    assert st.seg("mlp").numel() == st.off_bias + (per_head if args.n_levels_bias else 0)
    # segments: 16-byte aligned, trainable prefix first, frozen tensors behind it
    for name, sl in st.offsets.items():
        assert sl.start % 4 == 0, name
    names = list(st.offsets)
    assert names[:2] == ["table", "mlp"]
    if args.no_transformation_optimization:
        assert st.offsets["axisangle"].start >= st.n_train
    else:
        assert st.offsets["axisangle"].stop <= st.n_train
    assert st.cfg.n_levels_bias == args.n_levels_bias and st.cfg.w_bias == pytest.approx(args.weight_bias)
    # flat16 is the rounded flat
    assert torch.equal(st.flat16, st.flat.to(torch.float16))
    # heads: what the kernel reads is what the modules hold
    mlp = st.seg("mlp")
    d = model.inr.density_net.params.detach()
    assert torch.equal(mlp[: d.numel()], d)
    if not args.no_pixel_variance:
        w0 = model.sigma_net.weight_views()[0].detach()  # logical columns: [slice embedding (16) | z1..z15 | pad]
        p0 = mlp[st.off_sigma : st.off_sigma + W * 32].view(W, 32)
        assert torch.equal(p0[:, :16], w0[:, :16]) and torch.equal(p0[:, 17:32], w0[:, 16:31])
        assert (p0[:, 16] == 0).all()  # z0 never reaches sigma_net (models.py:352: z[..., 1:])
    if args.n_levels_bias:
        b = model.b_net.params.detach()
        assert b.numel() == per_head
        assert torch.equal(mlp[st.off_bias : st.off_bias + per_head], b)
    # round trip after an "optimiser step" on the flat buffer
    with torch.no_grad():
        st.flat[: st.n_train].mul_(1.5)
    expect = {n: p.detach().clone() for n, p in model.named_parameters()}
    st.push_to_model()
    seg_of = {"slice_embedding.weight": "slice_embedding", "logit_coef": "logit_coef", "log_var_slice": "log_var_slice", "axisangle": "axisangle"}
    for n, p in model.named_parameters():
        if n in seg_of and (seg_of[n] not in st.offsets or st.offsets[seg_of[n]].start >= st.n_train):
            assert torch.equal(p, expect[n]), n  # unused by this configuration: stays behind the trainable prefix, untouched
        elif n == "sigma_net.params":
            got, ref = p.detach().view(-1)[: W * 32].view(W, 32), expect[n].view(-1)[: W * 32].view(W, 32)
            assert torch.allclose(got[:, :31], ref[:, :31] * 1.5)
            assert (got[:, 31] == 0).all()  # the pad column multiplies zeros: dropped by the packed layout
        elif p.numel():
            assert torch.allclose(p.detach(), expect[n] * 1.5), n
    assert set(st.loss_dict(st.losses)) >= ({"MSE", "imageReg"} | ({"biasReg"} if args.n_levels_bias else set()))


def test_unsupported_configurations_raise(native_lib):
    from nesvor_b200.nesvor.fused import FusedState, FusedUnsupported

    model = build_model(make_args())
    for kw in (dict(n_levels_bias=5), dict(n_levels_bias=4, no_pixel_variance=True), dict(depth=4), dict(width=48),
               dict(n_features_z=7), dict(depth=2)):
        with pytest.raises(FusedUnsupported):
            FusedState(model.inr, make_args(**kw), model)


def test_locality_aware_batch_order_keeps_the_batches():
    """Dataset(locality_batch_size=B) (SURVEY s.8f row 1; reference batching: nesvor/nesvor/train.py:60-75): after the epoch
    shuffle every consecutive block of B pixels is put in (slice, y, x) order -- same batch membership as the reference's
    randperm blocks, spatially ordered inside."""
    from argparse import Namespace

    import torch

    from nesvor_b200.nesvor.train import Dataset
    from nesvor_b200.transform import RigidTransform

    class FakeSlice:
        def __init__(self, i, n=500):
            g = torch.Generator().manual_seed(i)
            self.xyz_masked_untransformed = torch.cat([torch.randint(0, 60, (n, 2), generator=g).float() - 30, torch.zeros(n, 1)], 1)
            self.v_masked = torch.rand(n, generator=g)
            self.transformation = RigidTransform(torch.eye(3, 4)[None])
            self.resolution_xyz = torch.tensor([1.0, 1.0, 3.0])

    ds = Dataset([FakeSlice(i) for i in range(6)], Namespace(locality_batch_size=256))
    idx = torch.randperm(3000, generator=torch.Generator().manual_seed(0))
    out = ds._order_inside_batches(idx, 256)
    for b in range(0, 3000, 256):
        assert sorted(idx[b : b + 256].tolist()) == sorted(out[b : b + 256].tolist())
        sl = ds.slice_idx[out[b : b + 256]]
        assert (sl[1:] >= sl[:-1]).all()
        same = sl[1:] == sl[:-1]
        y = ds.xyz[out[b : b + 256], 1]
        assert (y[1:][same] >= y[:-1][same]).all()
    # and through get_batch: an epoch of batches covers every pixel exactly once
    ds.count = ds.xyz.shape[0]
    seen = torch.cat([ds.get_batch(256, torch.device("cpu"))["v"] for _ in range(3000 // 256)])
    assert seen.numel() == 2816 and torch.unique(seen).numel() == seen.numel()


def test_loss_ring_slots_and_names(native_lib):
    """FusedState.next_losses: one zeroed 8-float slot per iteration, earlier slots keep their values until the ring wraps, the
    wrap clears the ring; loss_dict hands out views into the slot under the reference's names (no device work), "MSE+logVar" = word 6
    (written by the finalize kernel), transReg = word 5."""
    from nesvor_b200.nesvor import fused as F

    args = make_args()
    model = build_model(args)
    st = F.FusedState(model.inr, args, model)
    assert st.loss_ring.shape == (F.LOSS_RING, 8) and st.losses.data_ptr() == st.loss_ring.data_ptr()
    a = st.next_losses()
    a += torch.arange(8.0)
    names = st.loss_dict(a)
    assert float(names["MSE"]) == 0.0 and float(names["logVar"]) == 1.0 and float(names["MSE+logVar"]) == 6.0 and float(names["imageReg"]) == 3.0
    assert all(v.untyped_storage().data_ptr() == st.loss_ring.untyped_storage().data_ptr() for v in names.values())
    b = st.next_losses()
    assert b.data_ptr() == a.data_ptr() + 32 and not b.any() and float(a[6]) == 6.0  # a fresh slot; the previous one is intact
    st.loss_slot = F.LOSS_RING - 1
    last = st.losses = st.loss_ring[st.loss_slot]
    last += 1
    c = st.next_losses()  # wraps
    assert st.loss_slot == 0 and c.data_ptr() == st.loss_ring.data_ptr() and not st.loss_ring.any()


def test_host_batch_feeder_bookkeeping(monkeypatch):
    """HostBatchFeeder's slot accounting without a GPU (streams and events stubbed; the overlap itself is a GPU test,
    tests/test_gpu_e2e.py): batches come out in order, one copy ahead, every slot is handed back -- also when the consumer leaves the
    loop early --, and `bytes_copied` counts what was copied."""
    import contextlib

    from nesvor_b200.nesvor import fused as F

    log = []

    class Ev:
        def record(self, stream=None):
            log.append("record")

    class Stream:
        def wait_event(self, ev):
            log.append("wait")

    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None: Stream())
    monkeypatch.setattr(torch.cuda, "Event", lambda **kw: Ev())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: Stream())
    feeder = F.HostBatchFeeder("cpu")
    host = [{"v": torch.full((4,), float(i)), "slice_idx": torch.full((4,), i, dtype=torch.int64)} for i in range(5)]
    seen = []
    for batch in feeder.feed(host):
        assert feeder._head - feeder._tail in (1, 2)  # this batch + (except at the end) the next one already on its way
        seen.append((float(batch["v"][0]), int(batch["slice_idx"][0])))
    assert seen == [(float(i), i) for i in range(5)] and feeder._head == feeder._tail == 5
    assert feeder.bytes_copied == 5 * (16 + 32)
    for i, batch in enumerate(feeder.feed(host)):
        if i == 1:
            break  # two batches consumed, a third in flight
    assert feeder._head == feeder._tail  # the generator's clean-up released the outstanding slots
    assert [float(b["v"][0]) for b in feeder.feed(host[:3])] == [0.0, 1.0, 2.0]
    assert list(feeder.feed([])) == []
    with pytest.raises(RuntimeError):
        feeder.pop()
    feeder.push(host[0])
    feeder.push(host[1])
    with pytest.raises(RuntimeError):
        feeder.push(host[2])  # depth 2: both slots hold unreleased batches


@pytest.mark.parametrize("t_uni,t_mc,want", [(0.131, 0.105, True), (0.110, 0.111, False), (0.100, 0.098, False)])
def test_exchange_autotune_decision(native_lib, monkeypatch, t_uni, t_mc, want):
    """FusedTrainer._autotune_multimem without GPUs: both kernels are launched 2 + 8 times with lr = 0, gradient scale 0, step 1
    (state-preserving), the timings are max-reduced over the ranks, multicast is taken only when it wins by more than 3 %."""
    from nesvor_b200.nesvor import fused as F

    args = make_args()
    tr = F.FusedTrainer(build_model(args), args)
    tr.state.mc_ptrs = (1 << 40, (1 << 40) + 4096, 0)
    calls = []
    monkeypatch.setattr(tr, "_peer_exchange", lambda world, rank, mc, lr, unscale, step: calls.append((mc is not None, lr, unscale, step)))
    pending = [t_uni * 8, t_mc * 8]

    class Ev:
        def __init__(self, enable_timing=False):
            pass

        def record(self, stream=None):
            pass

        def elapsed_time(self, other):
            return pending.pop(0)

    monkeypatch.setattr(torch.cuda, "Event", Ev)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda device=None: None)

    class Dist:
        class ReduceOp:
            MAX = "max"

        reduced = []

        @staticmethod
        def all_reduce(t, op=None):
            Dist.reduced.append((t.clone(), op))

    assert tr._autotune_multimem(Dist, 8, 3) is want
    assert calls == [(False, 0.0, 0.0, 1)] * 10 + [(True, 0.0, 0.0, 1)] * 10
    assert Dist.reduced[-1][1] == "max" and Dist.reduced[-1][0].numel() == 2
    assert tr.dp_autotune_ms == pytest.approx({"unicast": t_uni, "multimem": t_mc}, rel=1e-5)
