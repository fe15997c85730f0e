"""CPU test of the inference callers (nesvor/nesvor/sample.py:10-64 mirrors): batching, masking and shapes of
`sample_points` / `sample_slice` / `sample_slices` / `sample_volume`, with the renderer replaced by a closed-form
stand-in (the renderer itself is covered on the GPU by tests/test_gpu_fused.py and tests/test_gpu_e2e.py)."""
from argparse import Namespace

import torch


class _FakeINR(torch.nn.Module):
    pass


def _args(**kw):
    a = dict(output_resolution=0.8, inference_batch_size=37, n_inference_samples=16, no_output_psf=False, device=torch.device("cpu"), fused=True)
    a.update(kw)
    return Namespace(**a)


def test_sample_functions_host_logic(monkeypatch):
    import nesvor_b200.nesvor.sample as sm
    from nesvor_b200.image import Slice, Volume
    from nesvor_b200.transform import RigidTransform, transform_points
    from nesvor_b200.utils import meshgrid, resolution2sigma

    calls = []

    def fake_render(model, xyz, transformation, psf_sigma, n_samples, args):
        calls.append(dict(n=xyz.shape[0], transformation=transformation, sigma=psf_sigma, n_samples=n_samples))
        world = xyz if transformation is None else transform_points(transformation, xyz)
        return (world * torch.tensor([1.0, 10.0, 100.0])).sum(-1).double()  # a dtype the callers must convert

    monkeypatch.setattr(sm, "_render", fake_render)
    model = _FakeINR()
    # ---- sample_points: any leading shape, chunks of inference_batch_size, float32 out
    xyz = torch.randn(5, 21, 3)
    v = sm.sample_points(model, xyz, _args())
    assert v.shape == (5, 21) and v.dtype == torch.float32
    assert torch.allclose(v, (xyz * torch.tensor([1.0, 10.0, 100.0])).sum(-1), atol=1e-4)
    assert [c["n"] for c in calls] == [37, 37, 31] and all(c["transformation"] is None and c["n_samples"] == 16 for c in calls)
    assert torch.allclose(torch.as_tensor(calls[0]["sigma"]), torch.as_tensor(resolution2sigma(0.8, isotropic=True)))
    calls.clear()
    sm.sample_points(model, xyz[:1, :3], _args(no_output_psf=True))
    assert calls[0]["n_samples"] == 0
    assert sm.sample_points(model, torch.zeros(0, 3), _args()).shape == (0,)
    # ---- sample_slice: only pixels whose world position falls inside the mask volume are rendered
    mvol = torch.zeros(9, 9, 9)
    mvol[2:7, 2:7, 2:7] = 1
    ident = torch.cat([torch.eye(3), torch.zeros(3, 1)], -1)[None]
    mask = Volume(mvol, mvol > 0, RigidTransform(ident, True), 1.0, 1.0, 1.0)
    pose = torch.tensor([[[1.0, 0, 0, 0.5], [0, 1.0, 0, -0.5], [0, 0, 1.0, 1.0]]])
    sl = Slice(torch.rand(1, 8, 8) + 1, torch.ones(1, 8, 8, dtype=torch.bool), RigidTransform(pose, True), 1.0, 1.0, 3.0)
    calls.clear()
    out = sm.sample_slice(model, sl, mask, _args())
    grid = meshgrid(sl.shape_xyz, sl.resolution_xyz).view(-1, 3)
    world = transform_points(sl.transformation, grid)
    inside = (mask.sample_points(world) > 0).view(1, 8, 8)
    assert 0 < int(inside.sum()) < 64 and torch.equal(out.mask, inside)
    expect = torch.zeros(1, 8, 8)
    expect[inside] = (world * torch.tensor([1.0, 10.0, 100.0])).sum(-1)[inside.view(-1)]
    assert torch.allclose(out.image, expect, atol=1e-4) and out.image.dtype == sl.image.dtype
    assert calls[0]["n"] == int(inside.sum()) and calls[0]["transformation"] is out.transformation
    assert torch.allclose(calls[0]["sigma"], resolution2sigma(sl.resolution_xyz, isotropic=False))
    assert torch.equal(sl.mask, torch.ones(1, 8, 8, dtype=torch.bool)) and float(sl.image.min()) >= 1  # the input slice is untouched
    # a slice entirely outside the mask: nothing rendered, empty mask
    far = Slice(torch.ones(1, 4, 4), None, RigidTransform(torch.tensor([[[1.0, 0, 0, 50.0], [0, 1.0, 0, 0], [0, 0, 1.0, 0]]]), True), 1.0, 1.0, 3.0)
    calls.clear()
    out = sm.sample_slice(model, far, mask, _args())
    assert not calls and not out.mask.any() and float(out.image.abs().sum()) == 0
    assert len(sm.sample_slices(model, [sl, far, sl], mask, _args())) == 3
    # ---- sample_volume: the mask resampled to the output resolution, filled inside its own mask
    calls.clear()
    vol = sm.sample_volume(model, mask, _args(output_resolution=0.5))
    assert isinstance(vol, Volume) and vol.resolution_x == 0.5 and int(vol.mask.sum()) == sum(c["n"] for c in calls) > 125
    w = vol.xyz_masked
    assert torch.allclose(vol.image[vol.mask], (w * torch.tensor([1.0, 10.0, 100.0])).sum(-1), atol=1e-3)
