"""Checkpoint round trip and reading a reference-style `model.pt` (nesvor/cli/io.py:36-59): pickled
`nesvor.image.image.Volume` / `nesvor.transform.transform.RigidTransform` objects and tcnn-layout flat parameters
(first MLP layer padded to 16 input columns).  Host logic only: no kernel runs."""
import sys
import types
from argparse import Namespace

import torch


def _args(**kw):
    a = dict(n_features_per_level=2, log2_hashmap_size=12, level_scale=2.0, coarsest_resolution=16.0, finest_resolution=8.0,
             depth=1, width=32, n_features_z=15, single_precision=False, dtype=torch.float16, device=torch.device("cpu"))
    a.update(kw)
    return Namespace(**a)


def test_save_load_round_trip(tmp_path):
    from nesvor_b200.image import Volume
    from nesvor_b200.io import load_model, save_model
    from nesvor_b200.nesvor.models import INR
    from nesvor_b200.transform import RigidTransform

    args = _args()
    bb = torch.tensor([[-30.0, -30.0, -30.0], [30.0, 30.0, 30.0]])
    inr = INR(bb, args)
    mask = Volume(torch.ones(4, 5, 6), torch.ones(4, 5, 6, dtype=torch.bool), RigidTransform(torch.zeros(1, 6)), 0.8, 0.8, 0.8)
    path = str(tmp_path / "model.pt")
    save_model(path, inr, mask, args)
    inr2, mask2, args2 = load_model(path, torch.device("cpu"))
    for k, v in inr.state_dict().items():
        assert torch.equal(v, inr2.state_dict()[k]), k
    assert torch.equal(mask2.mask, mask.mask) and mask2.resolution_x == 0.8 and args2.width == 32


def test_small_first_layer_is_written_in_the_16_padded_layout_and_round_trips(tmp_path):
    """2 levels x 2 features = 4 inputs: the kernels keep 32 input columns, tcnn pads to 16: the file carries the
    16-column first layer (64 x 16 + 64 x 16 = 2048 parameters here, not 64 x 32 + ...), and loading restores the module."""
    from nesvor_b200.image import Volume
    from nesvor_b200.io import load_model, save_model
    from nesvor_b200.nesvor.models import INR
    from nesvor_b200.transform import RigidTransform

    args = _args(width=64)
    inr = INR(torch.tensor([[-30.0, -30.0, -30.0], [30.0, 30.0, 30.0]]), args)
    assert inr.encoding.n_levels * 2 <= 16 and inr.density_net.layer_shapes[0] == (64, 32)
    with torch.no_grad():  # the padding columns multiply zeros: whatever they hold is immaterial and not stored
        inr.density_net.weight_views()[0][:, 16:] = 0
    mask = Volume(torch.ones(2, 2, 2), torch.ones(2, 2, 2, dtype=torch.bool), RigidTransform(torch.zeros(1, 6)), 1.0, 1.0, 1.0)
    path = str(tmp_path / "m.pt")
    save_model(path, inr, mask, args)
    raw = torch.load(path, weights_only=False)["model"]["density_net.params"]
    assert raw.numel() == 64 * 16 + 16 * 64
    inr2, _, _ = load_model(path, torch.device("cpu"))
    assert torch.equal(inr2.density_net.params, inr.density_net.params)


def test_load_builds_the_network_from_the_checkpoints_own_args(tmp_path):
    """cli/io.py:24-29: INR(cp["model"]["bounding_box"], cp["args"]) -- the caller's namespace, which `inputs()` passes in
    full, must not change the architecture the stored parameters are decoded with (a different coarsest_resolution with
    every level hashed would even keep the parameter count and silently decode on the wrong grid)."""
    from nesvor_b200.image import Volume
    from nesvor_b200.io import load_model, save_model
    from nesvor_b200.nesvor.models import INR
    from nesvor_b200.transform import RigidTransform

    args = _args()
    bb = torch.tensor([[-30.0, -30.0, -30.0], [30.0, 30.0, 30.0]])
    inr = INR(bb, args)
    mask = Volume(torch.ones(4, 5, 6), torch.ones(4, 5, 6, dtype=torch.bool), RigidTransform(torch.zeros(1, 6)), 0.8, 0.8, 0.8)
    path = str(tmp_path / "model.pt")
    save_model(path, inr, mask, args)
    caller = _args(width=64, depth=2, coarsest_resolution=8.0, finest_resolution=1.0, log2_hashmap_size=10, n_features_z=7,
                   inference_batch_size=123)
    inr2, _, merged = load_model(path, torch.device("cpu"), caller)
    for k, v in inr.state_dict().items():
        assert torch.equal(v, inr2.state_dict()[k]), k
    assert inr2.encoding.n_levels == inr.encoding.n_levels and inr2.encoding.meta.res[0] == inr.encoding.meta.res[0]
    assert merged.width == 64 and merged.inference_batch_size == 123  # merged namespace: the caller's values win (utils/misc.py:22-26)


def test_reads_reference_style_checkpoint(tmp_path):
    from nesvor_b200.io import load_model
    from nesvor_b200.nesvor.models import INR

    # stand-ins for the reference's classes, living under the reference's module paths while the file is written
    mods = {n: types.ModuleType(n) for n in ("nesvor", "nesvor.image", "nesvor.image.image", "nesvor.transform", "nesvor.transform.transform")}

    class RigidTransform:  # attribute names of nesvor/transform/transform.py:8-22
        def __init__(self, ax):
            self.trans_first, self._axisangle, self._matrix = True, ax, None

    class Volume:  # attribute names of nesvor/image/image.py:17-42
        def __init__(self, image, mask, transformation, r):
            self.image, self.mask, self.transformation = image, mask, transformation
            self.resolution_x = self.resolution_y = self.resolution_z = r

    RigidTransform.__module__, RigidTransform.__qualname__ = "nesvor.transform.transform", "RigidTransform"
    Volume.__module__, Volume.__qualname__ = "nesvor.image.image", "Volume"
    mods["nesvor.transform.transform"].RigidTransform = RigidTransform
    mods["nesvor.image.image"].Volume = Volume
    sys.modules.update(mods)
    try:
        args = _args()
        bb = torch.tensor([[-30.0, -30.0, -30.0], [30.0, 30.0, 30.0]])
        ours = INR(bb, args)  # 2 levels x 2 features = 4 inputs: tcnn pads the first layer to 16 columns, this build to 32
        n_table = ours.encoding.params.numel()
        g = torch.Generator().manual_seed(0)
        w0, w1 = torch.randn(32, 16, generator=g), torch.randn(16, 32, generator=g)
        state = {"bounding_box": bb, "encoding.params": torch.randn(n_table, generator=g),
                 "density_net.params": torch.cat([w0.reshape(-1), w1.reshape(-1)])}
        mask = Volume(torch.ones(3, 3, 3), torch.ones(3, 3, 3, dtype=torch.bool), RigidTransform(torch.zeros(1, 6)), 1.0)
        ref_args = Namespace(**{k: v for k, v in vars(args).items() if k not in ("dtype", "device")})
        path = str(tmp_path / "ref_model.pt")
        torch.save({"model": state, "mask": mask, "args": ref_args}, path)
    finally:
        for n in mods:
            sys.modules.pop(n, None)
    inr, mask2, merged = load_model(path, torch.device("cpu"))
    from nesvor_b200.image import Volume as OurVolume
    from nesvor_b200.transform import RigidTransform as OurRT

    assert isinstance(mask2, OurVolume) and isinstance(mask2.transformation, OurRT) and merged.dtype == torch.float16
    assert torch.equal(inr.encoding.params.detach(), state["encoding.params"])
    v0, v1 = inr.density_net.weight_views()
    assert torch.equal(v0[:, :16].detach(), w0) and float(v0[:, 16:].detach().abs().max()) == 0.0 and torch.equal(v1.detach(), w1)


def test_inputs_outputs_mirror_cli_io(tmp_path):
    """`outputs` / `inputs` (nesvor/cli/io.py:9-49): volume (rescaled), model, slice folders out; stacks (+ masks,
    thickness override), slices and model back in, with the checkpoint's args merged under the caller's."""
    import numpy as np

    from nesvor_b200.image import Volume
    from nesvor_b200.io import inputs, outputs
    from nesvor_b200.nesvor.models import INR
    from nesvor_b200.transform import RigidTransform

    rng = np.random.default_rng(0)
    args = _args()
    bb = torch.tensor([[-30.0, -30.0, -30.0], [30.0, 30.0, 30.0]])
    inr = INR(bb, args)
    img = torch.tensor(rng.uniform(0.2, 1.0, size=(4, 5, 6)), dtype=torch.float32)
    ident = torch.tensor([[[1.0, 0, 0, 1.5], [0, 1.0, 0, -2.0], [0, 0, 1.0, 0.5]]])
    vol = Volume(img.clone(), img > 0.4, RigidTransform(ident, True), 0.8, 0.8, 2.4)
    pv, pmask = str(tmp_path / "stack.nii.gz"), str(tmp_path / "stack_mask.nii.gz")
    vol.save(pv, masked=False)
    Volume(vol.mask.float(), None, vol.transformation, 0.8, 0.8, 2.4).save(pmask, masked=False)
    # ---- inputs: stacks with masks and a thickness override
    a_in = Namespace(input_stacks=[pv, pv], stack_masks=[pmask, pmask], thicknesses=[3.0, 4.0], device=torch.device("cpu"))
    data, a_out = inputs(a_in)
    assert len(data["input_stacks"]) == 2 and data["input_stacks"][1].thickness == 4.0 and data["input_stacks"][0].gap == 2.4000000953674316
    assert torch.equal(data["input_stacks"][0].mask[:, 0], vol.mask) and a_out is a_in
    # ---- outputs: everything named in args and present in data
    stack = data["input_stacks"][0]
    out_args = Namespace(**vars(args), output_volume=str(tmp_path / "out.nii.gz"), output_intensity_mean=700.0,
                         output_model=str(tmp_path / "model.pt"), output_slices=str(tmp_path / "slices"),
                         simulated_slices=str(tmp_path / "sim"))
    mean_before = float(vol.image[vol.mask].mean())
    outputs({"output_volume": vol, "output_model": inr, "mask": vol, "output_slices": stack[:], "simulated_slices": stack[:2]}, out_args)
    assert abs(float(vol.image[vol.mask].mean()) - 700.0) < 1e-2 and mean_before < 1.0  # rescaled in place like the reference
    assert sorted(__import__("os").listdir(out_args.output_slices)) == [f"{i}.nii.gz" for i in range(4)]
    assert len(__import__("os").listdir(out_args.simulated_slices)) == 2
    # ---- inputs again: slices + model; checkpoint args merged, caller's win
    a_in2 = Namespace(input_slices=out_args.output_slices, input_model=out_args.output_model, device=torch.device("cpu"), width=32, extra=5)
    data2, merged = inputs(a_in2)
    assert len(data2["input_slices"]) == 4 and isinstance(data2["mask"], Volume)
    for k, v in inr.state_dict().items():
        assert torch.equal(v, data2["model"].state_dict()[k]), k
    assert merged.extra == 5 and merged.output_intensity_mean == 700.0 and merged.n_features_z == 15
    sl = data2["input_slices"][2]
    np.testing.assert_allclose(sl.transformation.matrix(True).numpy(), stack[2].transformation.matrix(True).numpy(), atol=2e-4)
    from nesvor_b200.image import load_volume

    back = load_volume(out_args.output_volume)
    assert torch.allclose(back.image, vol.image * vol.mask, rtol=1e-6)


def test_reads_a_checkpoint_written_by_the_reference_writer(tmp_path):
    """tests/golden/reference_model.pt was written by the reference's OWN `cli/io.py::outputs` with the reference's own INR /
    Volume / RigidTransform classes (tests/golden/make_golden_model_pt.py, on nesvor_b200.compat): `load_model` must rebuild
    the INR from the file's args and return the stored tensors, mask and pose; when a copy of the reference is at hand the
    file is regenerated in a subprocess and must carry the same tensors (the fixture is not stale)."""
    import os
    import subprocess

    import numpy as np

    from nesvor_b200.io import load_model

    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    want = np.load(os.path.join(gold, "reference_model_tensors.npz"))

    def check(path):
        inr, mask, args = load_model(path, torch.device("cpu"))
        sd = inr.state_dict()
        for k in ("bounding_box", "encoding.params", "density_net.params"):
            assert np.array_equal(sd[k].float().numpy(), want[k]), k
        assert np.array_equal(mask.image.numpy(), want["mask_image"]) and np.array_equal(mask.mask.numpy(), want["mask_mask"])
        assert np.allclose(mask.transformation.axisangle().numpy(), want["mask_axisangle"], atol=1e-6)
        assert type(mask).__module__.startswith("nesvor_b200") and args.width == 64 and mask.resolution_x == 0.8

    check(os.path.join(gold, "reference_model.pt"))
    root = os.path.dirname(gold.rstrip(os.sep).rsplit(os.sep, 1)[0])
    if any(os.path.isdir(os.path.join(p, "nesvor")) for p in ("/root/reference", os.path.join(root, "baseline", "_ref"))):
        r = subprocess.run([sys.executable, os.path.join(gold, "make_golden_model_pt.py"), "--out", str(tmp_path)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-500:]
        check(str(tmp_path / "reference_model.pt"))
