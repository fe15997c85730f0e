"""CPU tests of the INR oracle itself (oracle/inr_oracle.py): properties that pin the restated
tcnn semantics independently of any kernel."""
import numpy as np
import torch

from oracle import inr_oracle as io


def test_level_geometry_matches_survey_numbers():
    # SURVEY s.8d: config 2 -> 7 dense + 9 hashed levels, 5 124 512 entries; defaults on a 110 mm box -> base 7, L 12
    m = io.grid_meta(16, 2, 19, 9, 1.3819)
    assert int(m.offset[-1]) == 5124512 and int(m.hashed.sum()) == 9 and int(m.res[0]) == 9 and int(m.res[-1]) == 1152
    m = io.grid_meta(12, 2, 19, 7, 1.3819)
    assert int(m.offset[-1]) == 2604424
    m = io.grid_meta(12, 2, 19, 17, 1.3819)
    assert int(m.offset[-1]) == 4054160


def test_dense_level_interpolates_lattice_values_exactly():
    """At x = (v - 0.5) / scale the +0.5 shift lands on vertex v: the encoding returns that table row."""
    m = io.grid_meta(1, 2, 19, 8, 2.0)
    res, s = int(m.res[0]), float(m.scale[0])
    g = torch.Generator().manual_seed(0)
    table = torch.randn(m.n_params, generator=g, dtype=torch.float64)
    v = torch.tensor([[1, 2, 3], [4, 0, 6], [2, 5, 1]], dtype=torch.float64)
    x = (v - 0.5) / s
    out = io.hashgrid_encode(x, table, m)
    idx = (v[:, 0] + v[:, 1] * res + v[:, 2] * res * res).long()
    torch.testing.assert_close(out, table.view(-1, 2)[idx])


def test_hash_matches_the_published_primes():
    m = io.grid_meta(1, 2, 4, 64, 2.0)  # 64^3 cells into 16 entries: hashed
    assert bool(m.hashed[0])
    table = torch.arange(m.n_params, dtype=torch.float64)
    v = torch.tensor([[5, 7, 11]], dtype=torch.float64)
    out = io.hashgrid_encode((v - 0.5) / float(m.scale[0]), table, m)
    h = (5 * 1) ^ ((7 * 2654435761) & 0xFFFFFFFF) ^ ((11 * 805459861) & 0xFFFFFFFF)
    torch.testing.assert_close(out[0], table.view(-1, 2)[h % 16])


def test_input_gradient_formula():
    """d enc / d x from autograd equals scale * sum over the other dims' weights * (feat[+1] - feat[0]) (App. A)."""
    m = io.grid_meta(3, 2, 8, 4, 1.7)
    g = torch.Generator().manual_seed(1)
    table = torch.randn(m.n_params, generator=g, dtype=torch.float64)
    x = torch.rand(64, 3, generator=g, dtype=torch.float64).requires_grad_(True)
    out = io.hashgrid_encode(x, table, m)
    go = torch.randn(out.shape, generator=g, dtype=torch.float64)
    (gx,) = torch.autograd.grad(out, x, go)
    eps = 1e-6
    for d in range(3):
        xp = x.detach().clone()
        xp[:, d] += eps
        fd = ((io.hashgrid_encode(xp, table, m) - out.detach()) * go).sum(-1) / eps
        torch.testing.assert_close(gx[:, d], fd, rtol=1e-4, atol=1e-5)


def test_fp16_emulation_is_a_small_perturbation():
    m = io.grid_meta(6, 2, 12, 5, 1.8)
    g = torch.Generator().manual_seed(2)
    table = torch.randn(m.n_params, generator=g) * 0.3
    x = torch.rand(2000, 3, generator=g)
    a, b = io.hashgrid_encode(x, table, m, False), io.hashgrid_encode(x, table, m, True)
    assert float((a - b).norm() / a.norm()) < 1e-3
    ws = [torch.randn(64, 12, generator=g) * 0.3, torch.randn(16, 64, generator=g) * 0.2]
    y, y16 = io.mlp_forward(a, ws), io.mlp_forward(a, ws, None, True)
    assert float((y - y16).norm() / y.norm()) < 3e-3


def test_analytic_loss_gradients_of_appendix_b():
    """SURVEY App. B (the formulas kernel A implements) vs autograd on the oracle's forward, fp64."""
    torch.manual_seed(0)
    B, S, ns = 7, 8, 3
    z0 = torch.randn(B, S, dtype=torch.float64, requires_grad=True)
    lv = (torch.randn(B, S, dtype=torch.float64) * 0.3).requires_grad_(True)
    lvs = (torch.randn(ns, dtype=torch.float64) * 0.3).requires_grad_(True)
    logit = (torch.randn(ns, dtype=torch.float64) * 0.3).requires_grad_(True)
    k = torch.randint(0, ns, (B,))
    v = torch.rand(B, dtype=torch.float64)
    x = torch.randn(B, S, 3, dtype=torch.float64)
    delta, w_i = 0.15, 2.0
    rho = torch.nn.functional.softplus(z0)
    c = torch.softmax(logit, 0)[k] * ns
    vhat = c * rho.mean(-1)
    var = (c.detach() * lv.exp().mean(-1)) ** 2 + lvs.exp()[k]
    loss = ((vhat - v) ** 2 / (2 * var)).mean() + 0.5 * var.log().mean() + w_i * io.image_reg("edge", rho, x, delta)
    gz0, glv, glvs, glogit = torch.autograd.grad(loss, (z0, lv, lvs, logit))
    with torch.no_grad():
        e = vhat - v
        d_vhat = e / (B * var)
        d_var = (0.5 / var - 0.5 * e * e / var**2) / B
        m = rho.mean(-1)
        r = c * lv.exp().mean(-1)
        dr = rho - rho.flip(1)
        d2 = ((x - x.flip(1)) ** 2).sum(-1) + 1e-6
        d_rho = (c * d_vhat / S)[:, None] + w_i * 2 * dr / (B * S * delta * d2 * torch.sqrt(1 + dr**2 / (d2 * delta**2)))
        torch.testing.assert_close(gz0, torch.sigmoid(z0) * d_rho)
        torch.testing.assert_close(glv, (lv.exp() / S) * (c * 2 * r * d_var)[:, None])
        torch.testing.assert_close(glvs, torch.zeros(ns, dtype=torch.float64).index_add_(0, k, lvs.exp()[k] * d_var))
        gc = torch.zeros(ns, dtype=torch.float64).index_add_(0, k, m * d_vhat)
        cs = torch.softmax(logit, 0) * ns
        torch.testing.assert_close(glogit, cs * (gc - (gc * cs).sum() / ns))


def test_analytic_gradients_of_the_bias_field_head():
    """The bias-field formulas kernel A implements (DESIGN s.4: v_out = c mean(exp(lb) rho), var through the DETACHED bias,
    biasReg = mean(lb)^2 with one batch-wide cotangent) vs autograd on the reference's op sequence (models.py:286-323), fp64."""
    torch.manual_seed(1)
    B, S, ns = 6, 8, 3
    z0 = torch.randn(B, S, dtype=torch.float64, requires_grad=True)
    lv = (torch.randn(B, S, dtype=torch.float64) * 0.3).requires_grad_(True)
    lb = (torch.randn(B, S, dtype=torch.float64) * 0.4).requires_grad_(True)
    lvs = (torch.randn(ns, dtype=torch.float64) * 0.3).requires_grad_(True)
    logit = (torch.randn(ns, dtype=torch.float64) * 0.3).requires_grad_(True)
    k = torch.randint(0, ns, (B,))
    v = torch.rand(B, dtype=torch.float64)
    w_b = 100.0
    rho = torch.nn.functional.softplus(z0)
    bias = lb.exp()
    c = torch.softmax(logit, 0)[k] * ns
    vhat = c * (bias * rho).mean(-1)
    var = (c.detach() * (bias.detach() * lv.exp()).mean(-1)) ** 2 + lvs.exp()[k]
    loss = ((vhat - v) ** 2 / (2 * var)).mean() + 0.5 * var.log().mean() + w_b * lb.mean() ** 2
    gz0, glv, glb, glogit = torch.autograd.grad(loss, (z0, lv, lb, logit))
    with torch.no_grad():
        e = vhat - v
        d_vhat = (e / (B * var))[:, None]
        d_var = ((0.5 / var - 0.5 * e * e / var**2) / B)[:, None]
        ck = c[:, None]
        r = ck * (bias * lv.exp()).mean(-1, keepdim=True)
        torch.testing.assert_close(gz0, torch.sigmoid(z0) * (ck * d_vhat / S * bias))
        torch.testing.assert_close(glv, (bias * lv.exp() / S) * ck * 2 * r * d_var)
        torch.testing.assert_close(glb, ck * d_vhat / S * bias * rho + w_b * 2 * lb.mean() / (B * S))
        m = (bias * rho).mean(-1)
        gc = torch.zeros(ns, dtype=torch.float64).index_add_(0, k, m * d_vhat[:, 0])
        cs = torch.softmax(logit, 0) * ns
        torch.testing.assert_close(glogit, cs * (gc - (gc * cs).sum() / ns))


def test_oracle_bias_head_wiring():
    """OracleNeSVoR with n_levels_bias: b_net sees [slice embedding | first n_levels_bias levels] (models.py:344-347) and
    nothing else -- perturbing finer levels leaves log_bias (hence biasReg) unchanged, perturbing the coarse ones does not;
    the gradient of biasReg alone reaches b_net, the slice embedding, the coarse levels and the poses only."""
    cfg = io.INRConfig(n_levels=6, base_resolution=5, level_scale=1.5, log2_hashmap_size=12, width=32, depth=1, n_levels_bias=2,
                       n_samples=8, no_transformation_optimization=False)
    g = torch.Generator().manual_seed(5)
    ns, B, S = 4, 12, 8
    ax = torch.randn(ns, 6, generator=g) * torch.tensor([0.1, 0.1, 0.1, 2.0, 2.0, 2.0])
    res = torch.tensor([[1.0, 1.0, 3.0]]).repeat(ns, 1)
    bb = torch.tensor([[-20.0, -20.0, -20.0], [20.0, 20.0, 20.0]])
    om = io.OracleNeSVoR(cfg, ns, ax, res, bb)
    with torch.no_grad():
        om.P["table"].copy_(torch.randn(om.P["table"].shape, generator=g) * 0.5)
    xyz = (torch.rand(B, 3, generator=g) - 0.5) * 20
    v = torch.rand(B, generator=g)
    idx = torch.randint(0, ns, (B,), generator=g)
    noise = torch.randn(B, S, 3, generator=g)
    l0 = om.forward(xyz, v, idx, noise)
    assert "biasReg" in l0 and float(l0["biasReg"].detach()) >= 0
    off = om.meta.offset
    with torch.no_grad():  # levels >= n_levels_bias do not feed b_net
        om.P["table"][2 * int(off[2]) :].mul_(1.7)
    l1 = om.forward(xyz, v, idx, noise)
    torch.testing.assert_close(l1["biasReg"], l0["biasReg"])
    assert not torch.allclose(l1["MSE"], l0["MSE"])
    with torch.no_grad():  # the coarse levels do
        om.P["table"][: 2 * int(off[2])].mul_(1.7)
    l2 = om.forward(xyz, v, idx, noise)
    assert not torch.allclose(l2["biasReg"], l1["biasReg"])
    # gradient of biasReg alone reaches b_net, the slice embedding, the coarse levels and the poses only
    (om.forward(xyz, v, idx, noise)["biasReg"]).backward()
    gt = om.P["table"].grad
    assert float(gt[: 2 * int(off[2])].abs().sum()) > 0 and float(gt[2 * int(off[2]) :].abs().sum()) == 0
    assert float(om.P["b_net.w0"].grad.abs().sum()) > 0 and float(om.P["slice_embedding"].grad.abs().sum()) > 0
    assert float(om.P["axisangle"].grad.abs().sum()) > 0
    assert om.P["density_net.w0"].grad is None or float(om.P["density_net.w0"].grad.abs().sum()) == 0
