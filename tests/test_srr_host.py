"""Host logic of nesvor_b200/svort/srr.py that needs no GPU: the CG solver on a dense SPD system and the regulariser
term dR against an independent per-voxel restatement of nesvor/svort/srr.py:139-160."""
import numpy as np
import torch


def test_cg_solves_spd_system_and_honours_stopping_rules():
    from nesvor_b200.svort.srr import CG

    g = torch.Generator().manual_seed(0)
    M = torch.randn(12, 12, generator=g, dtype=torch.float64)
    S = M @ M.T + 0.5 * torch.eye(12, dtype=torch.float64)
    x_true = torch.randn(12, generator=g, dtype=torch.float64)
    b = S @ x_true
    calls = []

    def A(x):
        calls.append(1)
        return S @ x

    x = CG(A, b, None, 12)
    torch.testing.assert_close(x, x_true, atol=1e-8, rtol=1e-8)
    assert len(calls) == 12  # x0 = None never evaluates A(0)
    calls.clear()
    x1 = CG(A, b, torch.zeros(12, dtype=torch.float64), 3)
    assert len(calls) == 1 + 3  # residual of x0 + one product per iteration, no extra product after the last update
    assert (S @ x1 - b).norm() < b.norm()
    calls.clear()
    x2 = CG(A, b, x_true + 1e-3, 50, tol=1e-20)  # started next to the solution: the tolerance stops it well before n_iter
    torch.testing.assert_close(x2, x_true, atol=1e-8, rtol=1e-8)
    assert len(calls) <= 1 + 12


def test_dR_matches_per_voxel_restatement():
    from nesvor_b200.svort.srr import SRR

    g = torch.Generator().manual_seed(1)
    v = torch.rand(1, 1, 5, 6, 7, generator=g, dtype=torch.float64)
    delta = 0.1
    got = SRR.dR(v, delta)[0, 0].numpy()
    a = v[0, 0].numpy()
    ref = np.zeros_like(a)
    D, H, W = a.shape
    for z in range(1, D - 1):
        for y in range(1, H - 1):
            for x in range(1, W - 1):
                acc = 0.0
                for dz in (-1, 0, 1):
                    for dy in (-1, 0, 1):
                        for dx in (-1, 0, 1):
                            if dx == dy == dz == 0:
                                continue
                            d = a[z, y, x] - a[z + dz, y + dy, x + dx]
                            s = d / (dx * dx + dy * dy + dz * dz) / (delta * delta)
                            acc += s / np.sqrt(1 + d * s)
                ref[z, y, x] = acc
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)
    assert got[0].max() == 0 and got[:, 0].max() == 0 and got[:, :, 0].max() == 0
