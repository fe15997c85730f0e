"""GPU parity: native pose converters (through the C ABI) vs the reference goldens and the oracle."""
import os

import numpy as np
import pytest
import torch

from helpers import REF_AXISANGLES, rel_l2, scipy_axisangle2mat

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_reference_golden_vectors(native_lib):
    import nesvor_b200 as nb

    for row in REF_AXISANGLES:  # one row at a time, like tests/transform/test_transform_convert.py:13-21
        ax = torch.tensor([row], dtype=torch.float32).cuda()
        mat = torch.from_numpy(scipy_axisangle2mat(ax.cpu().numpy())).cuda()
        torch.testing.assert_close(nb.axisangle2mat(ax), mat)
        torch.testing.assert_close(nb.mat2axisangle(mat), ax)


def test_against_reference_kernel_outputs(native_lib):
    """tests/golden/pose_ref.npz = the reference's own kernels run on CPU; tolerance = fp32 libm
    differences between glibc and CUDA (sinf/cosf/atan2f <= 2 ulp) amplified by |t| ~ 1."""
    import importlib

    tc = importlib.import_module("nesvor_b200.transform.transform_convert")

    g = {k: torch.from_numpy(v).cuda() for k, v in np.load(os.path.join(GOLD, "pose_ref.npz")).items()}
    torch.testing.assert_close(tc.axisangle2mat_forward(g["axisangle"])[0], g["mat"], atol=2e-6, rtol=1e-5)
    torch.testing.assert_close(tc.axisangle2mat_backward(g["grad_mat"], g["axisangle"])[0], g["a2m_bwd"], atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(tc.mat2axisangle_forward(g["mat"])[0], g["m2a_fwd"], atol=2e-5, rtol=1e-4)
    # near-pi rows amplify 1-ulp input differences; compare the bulk tightly and everything loosely
    torch.testing.assert_close(tc.mat2axisangle_backward(g["mat"], g["grad_axisangle"])[0], g["m2a_bwd"], atol=5e-2, rtol=1e-2)
    bulk = slice(11, None)
    torch.testing.assert_close(tc.mat2axisangle_backward(g["mat"][bulk].contiguous(), g["grad_axisangle"][bulk].contiguous())[0],
                               g["m2a_bwd"][bulk], atol=2e-4, rtol=1e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_against_oracle_random(native_lib, oracle, dtype):
    import importlib

    tc = importlib.import_module("nesvor_b200.transform.transform_convert")

    rng = np.random.default_rng(0)
    npdt = np.float32 if dtype == torch.float32 else np.float64
    ax = rng.normal(size=(4096, 6)).astype(npdt)
    ax[:64, :3] *= 1e-4
    axc = torch.from_numpy(ax).cuda()
    mat = tc.axisangle2mat_forward(axc)[0]
    torch.testing.assert_close(mat.cpu(), torch.from_numpy(oracle.axisangle2mat_forward(ax)[0]), atol=2e-6, rtol=1e-5)
    gm = rng.normal(size=(4096, 3, 4)).astype(npdt)
    torch.testing.assert_close(tc.axisangle2mat_backward(torch.from_numpy(gm).cuda(), axc)[0].cpu(),
                               torch.from_numpy(oracle.axisangle2mat_backward(gm, ax)[0]), atol=5e-5, rtol=1e-4)
    m_np = mat.cpu().numpy()
    torch.testing.assert_close(tc.mat2axisangle_forward(mat)[0].cpu(), torch.from_numpy(oracle.mat2axisangle_forward(m_np)[0]), atol=5e-5, rtol=1e-4)


def test_autograd_and_rigid_transform(native_lib):
    import nesvor_b200 as nb

    ax = torch.tensor(REF_AXISANGLES, dtype=torch.float32).cuda()
    zeros = torch.zeros(1, 6, device="cuda")
    for i in range(len(ax)):  # compose / inv identity, tests/transform/test_transform.py:7-23
        a, b = ax[i : i + 1].contiguous(), ax[len(ax) - 1 - i : len(ax) - i].contiguous()
        ma, mb = nb.axisangle2mat(a), nb.axisangle2mat(b)
        ab = nb.RigidTransform(a, trans_first=i % 2 == 0).compose(nb.RigidTransform(mb, trans_first=i % 2 == 1))
        binv_ainv = nb.RigidTransform(b, trans_first=i % 2 == 1).inv().compose(nb.RigidTransform(ma, trans_first=i % 2 == 0).inv())
        # the reference asserts atol 2e-5; translations here reach 300 (fp32 ulp 3e-5), so the bound is scaled by |t|
        scale = max(1.0, float(a[0, 3:].abs().max()), float(b[0, 3:].abs().max()))
        torch.testing.assert_close(ab.compose(binv_ainv).axisangle(), zeros, atol=2e-5 * scale, rtol=1e-3)
    x = torch.randn(16, 6, device="cuda", dtype=torch.float64, requires_grad=True)
    # analytic backward kernels vs numerical differentiation (trig is single precision inside)
    assert torch.autograd.gradcheck(nb.axisangle2mat, (x,), eps=1e-3, atol=1e-3, rtol=1e-3, nondet_tol=0)
    assert nb.axisangle2mat(torch.zeros(0, 6, device="cuda")).shape == (0, 3, 4)
    with pytest.raises(RuntimeError, match="contiguous"):
        nb.axisangle2mat(torch.zeros(6, 4, device="cuda").t())


def test_trans_reg_matches_autograd_composition_and_oracle(native_lib):
    """nsv_trans_reg_f32 (transReg + gradient in one launch; NeSVoR.trans_loss, models.py:357-363) vs (a) the same loss
    composed from RigidTransform.inv / compose / axisangle under autograd over the native converters and (b) the CPU
    oracle's trans_loss.  Cases: small and large relative motions, identical poses (zero error: the converters' first-order
    branches), large relative rotations.  fp32 tolerance 2e-5 relative on the loss, 2e-4 relative L2 on the gradient."""
    import ctypes

    import nesvor_b200 as nb
    from nesvor_b200 import _lib
    from oracle import inr_oracle as io

    g = torch.Generator().manual_seed(4)
    n = 300
    ax0 = torch.randn(n, 6, generator=g) * torch.tensor([0.8, 0.8, 0.8, 20.0, 20.0, 20.0])
    d = torch.randn(n, 6, generator=g) * torch.tensor([0.05, 0.05, 0.05, 1.0, 1.0, 1.0])
    d[:50] *= 10.0   # large relative motion (rotations of ~1 rad, translations of ~10 mm)
    d[50:60] = 0.0   # identical poses
    d[60:70] *= 1e-4  # nearly identical
    ax = ax0 + d
    for weight in (1.0, 0.1):
        a_dev, a0_dev = ax.cuda().contiguous(), ax0.cuda().contiguous()
        grad = torch.full((n, 6), 0.5, device="cuda")  # the kernel accumulates
        loss = torch.full((1,), 2.0, device="cuda")
        rc = _lib.lib().nsv_trans_reg_f32(_lib.ptr(a_dev), _lib.ptr(a0_dev), _lib.ptr(grad), _lib.ptr(loss), ctypes.c_int(n),
                                          ctypes.c_float(weight), _lib.stream(a_dev.device))
        _lib.check(rc, "nsv_trans_reg_f32")
        torch.cuda.synchronize()
        # (a) autograd over the native converters
        x_ax = a_dev.clone().requires_grad_(True)
        x, y = nb.RigidTransform(x_ax, True), nb.RigidTransform(a0_dev, True)
        err = y.inv().compose(x).axisangle(True)
        ref = torch.mean(err[:, :3] ** 2) + 1e-3 * torch.mean(err[:, 3:] ** 2)
        (gref,) = torch.autograd.grad(ref, x_ax)
        assert abs(float(loss) - 2.0 - float(ref)) <= 2e-5 * abs(float(ref))
        assert rel_l2((grad - 0.5).cpu(), (weight * gref).cpu()) < 2e-4
        # (b) CPU oracle
        xo = ax.clone().requires_grad_(True)
        e = io.mat2axisangle(io.mat_compose(io.mat_inv(io.axisangle2mat(ax0)), io.axisangle2mat(xo)))
        lo = torch.mean(e[:, :3] ** 2) + 1e-3 * torch.mean(e[:, 3:] ** 2)
        (go,) = torch.autograd.grad(lo, xo)
        assert abs(float(loss) - 2.0 - float(lo)) <= 5e-5 * abs(float(lo))
        assert rel_l2((grad - 0.5).cpu(), weight * go) < 1e-3
    assert _lib.lib().nsv_trans_reg_f32(None, None, None, None, ctypes.c_int(0), ctypes.c_float(1.0), None) == 0
