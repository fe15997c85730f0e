"""CPU tests of the kernel-B consumers next to the path (nesvor/svort/inference.py:370-444, utils/loss.py): `ncc_loss`
against golden outputs of the reference's own function (tests/golden/make_golden_affine.py), and the host logic of
`reconstruct_from_stacks` / `simulated_ncc` with the native operators replaced by recording stand-ins (the operators
themselves are covered on the GPU by tests/test_gpu_slice_acq.py and tests/test_srr_host.py's GPU counterpart)."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ncc_ref.npz")


def test_ncc_loss_matches_reference_goldens():
    from nesvor_b200.utils import ncc_loss

    g = np.load(GOLD)
    I, J, mask = torch.tensor(g["I"]), torch.tensor(g["J"]), torch.tensor(g["mask"])
    V, W = torch.tensor(g["V"]), torch.tensor(g["W"])
    cases = {"global_masked": ncc_loss(I, J, mask, win=None, reduction="none"), "global_plain": ncc_loss(I, J, None, win=None, reduction="none"),
             "win9": ncc_loss(I, J, None, win=9), "win9_level1_masked": ncc_loss(I, J, mask, win=9, level=1),
             "win5_3d_mean": ncc_loss(V, W, None, win=5, reduction="mean"), "win5_3d_sum": ncc_loss(V, W, None, win=5, reduction="sum")}
    for k, v in cases.items():
        assert v.shape == g[k].shape, k
        np.testing.assert_allclose(v.numpy(), g[k], rtol=1e-5, atol=1e-6, err_msg=k)
    # identical images inside the mask correlate perfectly; the value is minus the squared coefficient
    assert torch.allclose(ncc_loss(I, I, mask, win=None), -torch.ones(5, 1), atol=1e-3)


def test_reconstruct_from_stacks_host_logic(monkeypatch):
    import nesvor_b200.svort.inference as inf
    from nesvor_b200.transform import RigidTransform

    calls = {}

    def fake_psfrec(mat, slices, slices_mask, vol_mask, params):
        calls["psfrec"] = dict(mat=mat.clone(), slices=slices.clone(), params=dict(params), masks=(slices_mask, vol_mask))
        return torch.full((1, 1) + tuple(params["volume_shape"]), 2.0)

    class FakeSRR:
        def __init__(self, n_iter, use_CG):
            calls["srr_init"] = (n_iter, use_CG)

        def __call__(self, mat, slices, volume, params, slices_mask=None):
            calls["srr"] = dict(mat=mat, volume=volume, slices_mask=slices_mask.clone(), slices=slices)
            return volume + 1

    monkeypatch.setattr(inf, "PSFreconstruction", fake_psfrec)
    monkeypatch.setattr(inf, "SRR", FakeSRR)
    monkeypatch.setattr(inf, "get_PSF", lambda res_ratio, device: torch.tensor(res_ratio))
    g = torch.Generator().manual_seed(0)
    stacks = [torch.rand(3, 1, 5, 8, generator=g) - 0.2, torch.rand(2, 1, 8, 6, generator=g) - 0.2, torch.rand(4, 1, 8, 8, generator=g)]
    mats = [torch.cat([torch.eye(3).expand(s.shape[0], 3, 3), torch.full((s.shape[0], 3, 1), 4.0 * (j + 1))], -1) for j, s in enumerate(stacks)]
    transforms = [RigidTransform(m, True) for m in mats]
    vol = inf.reconstruct_from_stacks(transforms, stacks, res_s=1.0, s_thick=2.5, res_r=0.8, n_stack_recon=2, volume_shape=(6, 7, 8))
    assert torch.equal(vol, torch.full((1, 1, 6, 7, 8), 3.0)) and calls["srr_init"] == (1, True)
    p = calls["psfrec"]
    assert p["slices"].shape == (5, 1, 8, 8)  # the first two stacks, padded to the largest in-plane size of ALL stacks
    assert torch.equal(p["slices"][:3, :, 1:6, :], stacks[0]) and float(p["slices"][:3, :, 0].abs().sum()) == 0  # rows 5 -> 8: 1 before, 2 after
    assert float(p["slices"][:3, :, 6:].abs().sum()) == 0
    assert torch.equal(p["slices"][3:, :, :, 1:7], stacks[1])  # columns 6 -> 8: 1 before, 1 after
    assert p["masks"] == (None, None) and p["params"]["slice_shape"] == (8, 8) and p["params"]["interp_psf"] is False
    assert torch.allclose(p["params"]["psf"], torch.tensor([1.25, 1.25, 3.125])) and p["params"]["volume_shape"] == (6, 7, 8)
    # translations in voxel units of the reconstruction grid, rotations untouched
    assert torch.allclose(p["mat"][:3, :, 3], torch.full((3, 3), 4.0 / 0.8)) and torch.allclose(p["mat"][3:, :, 3], torch.full((2, 3), 8.0 / 0.8))
    assert torch.equal(p["mat"][:, :, :3], torch.eye(3).expand(5, 3, 3))
    assert torch.equal(calls["srr"]["slices_mask"], p["slices"] > 0)
    # n_stack_recon = None uses every stack; the default grid is SVoRT's 256^3
    monkeypatch.setattr(inf, "PSFreconstruction", lambda mat, slices, a, b, params: calls.update(n=slices.shape[0], shape=params["volume_shape"]) or torch.zeros(1))
    inf.reconstruct_from_stacks(transforms, stacks, 1.0, 2.5, 0.8, None)
    assert calls["n"] == 9 and calls["shape"] == (256, 256, 256)


def test_simulated_ncc_host_logic(monkeypatch):
    import nesvor_b200.svort.inference as inf
    from nesvor_b200.transform import RigidTransform

    seen = []

    def fake_acq(mat, vol, vol_mask, slices_mask, psf, slice_shape, res_slice, need_weight, interp_psf):
        seen.append(dict(mat=mat, slices_mask=slices_mask, slice_shape=tuple(slice_shape), res_slice=res_slice, flags=(vol_mask, need_weight, interp_psf)))
        return vol  # the "simulated" stack handed in below

    monkeypatch.setattr(inf, "slice_acquisition", fake_acq)
    monkeypatch.setattr(inf, "get_PSF", lambda res_ratio, device: torch.tensor(res_ratio))
    g = torch.Generator().manual_seed(1)
    stack = torch.rand(4, 1, 6, 6, generator=g) - 0.3
    tr = RigidTransform(torch.cat([torch.eye(3).expand(4, 3, 3), torch.ones(4, 3, 1)], -1), True)
    ncc, w = inf.simulated_ncc([tr], [stack], stack.clone(), res_s=1.0, s_thick=3.0, res_r=0.5)
    assert ncc.shape == (4, 1) and w.shape == (4, 1)
    assert torch.allclose(ncc, -torch.ones(4, 1), atol=5e-3)  # a stack compared with itself inside its mask (eps in the denominator)
    assert torch.equal(w[:, 0], (stack > 0).sum((1, 2, 3)))
    s = seen[0]
    assert s["slice_shape"] == (6, 6) and s["res_slice"] == 2.0 and s["flags"] == (None, False, False)
    assert torch.equal(s["slices_mask"], stack > 0) and torch.allclose(s["mat"][:, :, 3], torch.full((4, 3), 2.0))
