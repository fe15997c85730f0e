"""CPU tests (no GPU): the C restatement of the pose converters vs (a) the reference's golden
vectors (scipy rotvec pairs, /root/reference/tests/__init__.py:17-38 and
tests/transform/test_transform_convert.py:13-21), (b) the reference's own kernels compiled for CPU
(oracle/_ref, bit-exact), (c) the differentiable torch restatement used by the INR oracle."""
import numpy as np
import pytest
import torch

from helpers import REF_AXISANGLES, scipy_axisangle2mat


def test_axisangle2mat_golden(oracle):
    ax = np.array(REF_AXISANGLES, np.float32)
    torch.testing.assert_close(torch.from_numpy(oracle.axisangle2mat_forward(ax)[0]), torch.from_numpy(scipy_axisangle2mat(ax)))


def test_mat2axisangle_golden(oracle):
    ax = np.array(REF_AXISANGLES, np.float32)
    mat = scipy_axisangle2mat(ax)
    torch.testing.assert_close(torch.from_numpy(oracle.mat2axisangle_forward(mat)[0]), torch.from_numpy(ax))


def _random_poses(dtype, n=200, seed=3):
    rng = np.random.default_rng(seed)
    ax = rng.normal(size=(n, 6)).astype(dtype)
    ax[:10, :3] *= 1e-4  # small-angle branch
    ax[10:20, :3] *= 2.9 / np.linalg.norm(ax[10:20, :3], axis=1, keepdims=True)  # near pi: non-trace quaternion branches
    return ax


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_bit_exact_vs_reference_kernels(oracle, reference_cpu, dtype):
    ax = _random_poses(dtype)
    rng = np.random.default_rng(4)
    m_o, m_r = oracle.axisangle2mat_forward(ax)[0], reference_cpu.axisangle2mat_forward(ax)[0]
    assert np.array_equal(m_o, m_r)
    g = rng.normal(size=m_o.shape).astype(dtype)
    assert np.array_equal(oracle.axisangle2mat_backward(g, ax)[0], reference_cpu.axisangle2mat_backward(g, ax)[0])
    assert np.array_equal(oracle.mat2axisangle_forward(m_o)[0], reference_cpu.mat2axisangle_forward(m_o)[0])
    ga = rng.normal(size=ax.shape).astype(dtype)
    assert np.array_equal(oracle.mat2axisangle_backward(m_o, ga)[0], reference_cpu.mat2axisangle_backward(m_o, ga)[0])


def test_backward_matches_autograd_of_torch_restatement(oracle):
    """The analytic VJPs (untested in the reference, SURVEY s.4) agree with autograd through the
    differentiable torch restatement, in fp64 (trig is single precision in the C code -> 1e-6)."""
    from oracle import inr_oracle as io

    ax = torch.from_numpy(_random_poses(np.float64, n=60)[20:])  # generic angles
    ax.requires_grad_(True)
    mat = io.axisangle2mat(ax)
    g = torch.randn_like(mat)
    (ga,) = torch.autograd.grad(mat, ax, g)
    ga_c = oracle.axisangle2mat_backward(g.numpy(), ax.detach().numpy())[0]
    torch.testing.assert_close(torch.from_numpy(ga_c), ga, atol=5e-6, rtol=1e-5)
    m = mat.detach().clone().requires_grad_(True)
    a2 = io.mat2axisangle(m)
    g2 = torch.randn_like(a2)
    (gm,) = torch.autograd.grad(a2, m, g2)
    gm_c = oracle.mat2axisangle_backward(m.detach().numpy(), g2.numpy())[0]
    torch.testing.assert_close(torch.from_numpy(gm_c), gm, atol=5e-6, rtol=1e-5)


def test_compose_inv_identity(oracle):
    """(ab)(b^-1 a^-1) = I, the reference's test_transform.py:7-23, on the torch restatement."""
    from oracle import inr_oracle as io

    ax = torch.tensor(REF_AXISANGLES, dtype=torch.float32)
    for i in range(len(ax)):
        a, b = io.axisangle2mat(ax[i : i + 1]), io.axisangle2mat(ax[-i - 1 : len(ax) - i])
        ab = io.mat_compose(a, b)
        binv_ainv = io.mat_compose(io.mat_inv(b), io.mat_inv(a))
        err = io.mat2axisangle(io.mat_compose(ab, binv_ainv))
        torch.testing.assert_close(err, torch.zeros_like(err), atol=2e-5 * max(1.0, float(ax[i, 3:].abs().max())), rtol=1e-3)
