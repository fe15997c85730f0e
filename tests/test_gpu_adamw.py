"""nsv_adamw_step (fused AdamW over the flat parameter vector) against torch.optim.AdamW as the reference configures it
(nesvor/nesvor/train.py:134-152: lr 5e-3, betas (0.9, 0.99), eps 1e-15, weight decay 1e-2), through the C ABI."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 3, 4, 1027, 65536 + 5])
def test_adamw_matches_torch(native_lib, n):
    from nesvor_b200 import _lib

    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(n)
    p0 = torch.randn(n, generator=g).to(dev)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=5e-3, betas=(0.9, 0.99), eps=1e-15, weight_decay=1e-2)
    p, m, v = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    p16 = torch.zeros(n, dtype=torch.float16, device=dev)
    scale = 128.0  # loss scale folded into the gradient, removed by grad_unscale
    for step in range(1, 6):
        grad = (torch.randn(n, generator=g) * (torch.rand(n, generator=g) > 0.3)).to(dev)  # some exact zeros, like untouched table entries
        ref.grad = grad.clone()
        opt.step()
        gbuf = (grad * scale).contiguous()
        rc = _lib.lib().nsv_adamw_step(_lib.ptr(p), _lib.ptr(gbuf), _lib.ptr(m), _lib.ptr(v), _lib.ptr(p16), ctypes.c_int64(n),
                                       ctypes.c_float(5e-3), ctypes.c_float(0.9), ctypes.c_float(0.99), ctypes.c_float(1e-15), ctypes.c_float(1e-2),
                                       ctypes.c_int(step), ctypes.c_float(1.0 / scale), ctypes.c_int(1), _lib.stream(dev))
        _lib.check(rc, "nsv_adamw_step")
        torch.cuda.synchronize()
        assert torch.count_nonzero(gbuf) == 0  # zero_grad = 1 clears the gradient for the next iteration
        torch.testing.assert_close(p, ref.detach(), rtol=2e-6, atol=2e-7)
        assert torch.equal(p16, p.half())


def test_adamw_skips_non_finite_gradient_elements(native_lib):
    """Found-inf guard (the reference's GradScaler skips a step with inf / NaN gradients, train.py:161-164,195): an element
    whose gradient is not finite keeps parameter, both moments and its fp16 copy; its neighbours are updated as usual; the
    gradient is cleared either way."""
    import ctypes

    from nesvor_b200 import _lib

    n = 4099
    g = torch.Generator().manual_seed(0)
    p = torch.randn(n, generator=g).cuda()
    grad = torch.randn(n, generator=g).cuda()
    m, v = torch.rand(n, generator=g).cuda() * 0.1, torch.rand(n, generator=g).cuda() * 0.01
    p16 = p.half()
    bad = torch.tensor([0, 5, 4097, 4098])
    grad[bad] = torch.tensor([float("inf"), float("nan"), float("-inf"), float("nan")]).cuda()
    p0, m0, v0, g0 = p.clone(), m.clone(), v.clone(), grad.clone()
    _lib.check(native_lib.nsv_adamw_step(_lib.ptr(p), _lib.ptr(grad), _lib.ptr(m), _lib.ptr(v), _lib.ptr(p16), ctypes.c_int64(n), ctypes.c_float(5e-3),
                                         ctypes.c_float(0.9), ctypes.c_float(0.99), ctypes.c_float(1e-15), ctypes.c_float(1e-2), ctypes.c_int(3),
                                         ctypes.c_float(1.0), ctypes.c_int(1), _lib.stream(p.device)))
    torch.cuda.synchronize()
    assert torch.isfinite(p).all() and torch.isfinite(m).all() and torch.isfinite(v).all() and (grad == 0).all()
    assert torch.equal(p[bad.cuda()], p0[bad.cuda()]) and torch.equal(m[bad.cuda()], m0[bad.cuda()]) and torch.equal(v[bad.cuda()], v0[bad.cuda()])
    ok = torch.ones(n, dtype=torch.bool)
    ok[bad] = False
    ok = ok.cuda()
    assert (p[ok] != p0[ok]).all() and torch.allclose(m[ok], 0.9 * m0[ok] + 0.1 * g0[ok], rtol=1e-6, atol=1e-7)
    assert torch.equal(p16, p.half())
