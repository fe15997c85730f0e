"""GPU parity for the standalone hash-grid encoding and fused-MLP ops (C ABI) vs the torch oracle."""
import numpy as np
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu


def _grid(L, F, T, base, s, dtype):
    from nesvor_b200.nesvor.encoding import HashGridEncoding

    enc = HashGridEncoding(3, dict(otype="HashGrid", n_levels=L, n_features_per_level=F, log2_hashmap_size=T, base_resolution=base,
                                   per_level_scale=s), dtype=dtype).cuda()
    with torch.no_grad():  # larger than the 1e-4 init so that errors are visible
        enc.params.copy_(torch.randn_like(enc.params) * 0.5)
    return enc


@pytest.mark.parametrize("cfg", [(16, 2, 19, 9, 1.3819), (12, 2, 19, 7, 1.3819), (2, 2, 19, 5, 2.0), (6, 4, 12, 5, 1.8), (5, 1, 10, 3, 2.0), (4, 8, 11, 4, 1.7)])
def test_hashgrid_fp32_forward_backward(native_lib, cfg):
    from oracle import inr_oracle as io

    L, F, T, base, s = cfg
    enc = _grid(L, F, T, base, s, torch.float32)
    meta = io.grid_meta(L, F, T, base, s)
    g = torch.Generator().manual_seed(1)
    N = 20000
    x = torch.rand(N, 3, generator=g)
    x[:50] = torch.rand(50, 3, generator=g) * 1.2 - 0.1  # slightly outside [0,1]: wrap-around indices
    x[50] = torch.tensor([1.0, 1.0, 1.0])
    x[51] = torch.tensor([0.0, 0.0, 0.0])
    go = torch.randn(N, L * F, generator=g)
    xo = x.clone().requires_grad_(True)
    tab = enc.params.detach().cpu().clone().requires_grad_(True)
    out_o = io.hashgrid_encode(xo, tab, meta)
    gx_o, gt_o = torch.autograd.grad(out_o, (xo, tab), go)
    xc = x.cuda().requires_grad_(True)
    out = enc(xc)
    gx, gt = torch.autograd.grad(out, (xc, enc.params), go.cuda())
    torch.testing.assert_close(out.cpu(), out_o.detach(), atol=2e-6, rtol=1e-5)
    assert rel_l2(gt.cpu(), gt_o) < 1e-5
    assert rel_l2(gx.cpu(), gx_o) < 1e-5


def test_hashgrid_fp16_mode(native_lib):
    from oracle import inr_oracle as io

    L, F, T, base, s = 16, 2, 19, 9, 1.3819
    enc = _grid(L, F, T, base, s, torch.float16)
    meta = io.grid_meta(L, F, T, base, s)
    x = torch.rand(50000, 3, generator=torch.Generator().manual_seed(2))
    out_o = io.hashgrid_encode(x, enc.params.detach().cpu(), meta, emulate_fp16=True)
    out = enc(x.cuda())
    assert out.dtype == torch.float16
    # identical rounding points (fp16 table, fp32 blend, fp16 output): at most 1 fp16 ulp apart
    assert rel_l2(out.float().cpu(), out_o) < 3e-4
    xc = x.cuda().requires_grad_(True)
    go = torch.randn(50000, L * F, device="cuda") * 1e-3
    gx, gt = torch.autograd.grad(enc(xc), (xc, enc.params), go.half())
    xo = x.clone().requires_grad_(True)
    tab = enc.params.detach().cpu().clone().requires_grad_(True)
    gx_o, gt_o = torch.autograd.grad(io.hashgrid_encode(xo, tab, meta, emulate_fp16=True), (xo, tab), go.half().float().cpu())
    assert rel_l2(gt.cpu(), gt_o) < 2e-3
    assert rel_l2(gx.cpu(), gx_o) < 2e-3


@pytest.mark.parametrize("shape", [(32, 16, 64, 3), (24, 16, 64, 1), (31, 1, 64, 1), (32, 16, 32, 2), (64, 16, 64, 2), (48, 5, 32, 1)])
def test_fused_mlp_forward_backward(native_lib, shape):
    """fp16 operands / fp32 accumulate vs the oracle with the same rounding points."""
    from nesvor_b200.nesvor.encoding import FusedMLP
    from oracle import inr_oracle as io

    n_in, n_out, width, depth = shape
    mlp = FusedMLP(n_in, n_out, dict(otype="CutlassMLP", activation="ReLU", output_activation="None", n_neurons=width, n_hidden_layers=depth)).cuda()
    g = torch.Generator().manual_seed(3)
    N = 1000  # not a multiple of the 256-row tile
    x = torch.randn(N, n_in, generator=g)
    ws = [w.detach().cpu().clone().requires_grad_(True) for w in mlp.weight_views()]
    xo = x.half().float().clone().requires_grad_(True)
    out_o = io.mlp_forward(xo, ws, None, emulate_fp16=True)[:, :n_out]
    xc = x.cuda().requires_grad_(True)
    out = mlp(xc)
    assert out.dtype == torch.float16 and out.shape == (N, n_out)
    assert rel_l2(out.float().cpu(), out_o.detach()) < 1e-3  # output itself is rounded to fp16
    go = torch.randn(N, n_out, generator=g) * 0.01
    gx, gw = torch.autograd.grad(out, (xc, mlp.params), go.cuda().half())
    grads_o = torch.autograd.grad(out_o, [xo] + ws, go.half().float())
    gw_o = torch.cat([t.reshape(-1) for t in grads_o[1:]])
    assert rel_l2(gw.cpu(), gw_o) < 5e-3
    assert rel_l2(gx.cpu(), grads_o[0]) < 5e-3
