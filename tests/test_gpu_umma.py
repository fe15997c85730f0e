"""GPU check of the tcgen05 / TMEM layer (nesvor_b200/csrc/umma.cuh): every operand configuration
kernel A uses, against torch matmuls in fp32 on the same fp16 inputs."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_umma_operand_layouts(native_lib):
    from nesvor_b200 import _lib

    g = torch.Generator().manual_seed(0)
    A = (torch.randn(128, 64, generator=g) * 0.5).half().cuda()
    W = (torch.randn(64, 64, generator=g) * 0.5).half().cuda()
    G = (torch.randn(128, 16, generator=g) * 0.5).half().cuda()
    sizes = [128 * 64, 128 * 64, 64 * 64, 64 * 64, 128 * 16, 128 * 64, 64 * 16, 128 * 64, 128 * 32, 64 * 32]
    out = torch.full((sum(sizes),), float("nan"), device="cuda")
    rc = _lib.lib().nsv_umma_selftest(_lib.ptr(A), _lib.ptr(W), _lib.ptr(G), _lib.ptr(out), _lib.stream())
    _lib.check(rc, "nsv_umma_selftest")
    torch.cuda.synchronize()
    parts = torch.split(out.cpu(), sizes)
    a, w, gg = A.float().cpu(), W.float().cpu(), G.float().cpu()
    expect = {
        "T1 fwd A W^T": a @ w.t(),
        "T2 dgrad A W": a @ w,
        "T3a wgrad A^T A": a.t() @ a,
        "T3b wgrad x2 @lane16": 2 * (a.t() @ a),
        "T4 out A Wo^T": a @ w[:16].t(),
        "T5 dgrad-out G Wo": gg @ w[:16],
        "T6 wgrad-out A^T G": a.t() @ gg,
        "T7 fwd K32": a[:, :32] @ w[:, :32].t(),
        "T8 dgrad N32": a @ w[:, :32],
        "T9 wgrad N32": a.t() @ a[:, :32],
    }
    bad = []
    for (name, ref), got in zip(expect.items(), parts):
        got = got.view_as(ref)
        err = float((got - ref).abs().max())
        print(f"{name}: max abs err {err:.3e} (ref max {float(ref.abs().max()):.2f})")
        if not (err < 2e-3 * max(1.0, float(ref.abs().max()))):
            bad.append(name)
    assert not bad, bad


@pytest.mark.parametrize("n_issuers,reps", [(1, 1), (1, 50), (2, 200), (3, 400), (4, 1000)])
def test_accumulating_mma_of_different_issuers_compose(native_lib, n_issuers, reps):
    """Hardware-behaviour probe the 3-group training kernel rests on: its three groups' issuing threads accumulate their
    weight-gradient products (M64 N64 K128, accumulate on) into ONE set of TMEM accumulators, unordered.  With exactly
    representable addends (small integers) every order of accumulation gives the same fp32 sum, so the result must be
    EXACTLY n_issuers * reps * A^T A -- a lost or torn update would show as a smaller count."""
    from nesvor_b200 import _lib

    g = torch.Generator().manual_seed(n_issuers * 1000 + reps)
    A = torch.randint(-2, 3, (128, 64), generator=g).half().cuda()  # integers: every partial sum stays exact in fp32 (< 2^24)
    out = torch.full((64, 64), float("nan"), device="cuda")
    for _ in range(5):  # a few launches: the interleaving of the issuers changes from run to run
        out.fill_(float("nan"))
        _lib.check(native_lib.nsv_umma_shared_accumulator_test(_lib.ptr(A), _lib.ptr(out), ctypes.c_int(n_issuers), ctypes.c_int(reps), _lib.stream()))
        torch.cuda.synchronize()
        want = (A.float().t() @ A.float()) * (n_issuers * reps)
        assert float(want.abs().max()) < 2**24
        assert torch.equal(out, want), (n_issuers, reps, float((out - want).abs().max()))
