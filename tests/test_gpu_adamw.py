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
