"""M2: phantom PSNR of the fused B200 training path vs the CPU oracle trained on the identical
batch / PSF-noise sequence from identical initial parameters.  Target (BASELINE.json north star): within 0.1 dB.
Short run of BASELINE config 1 (64^3 phantom, 3 stacks, 2-level hash grid, 32-wide MLP); the full 200-iteration
numbers and the config-2 model are produced by tools/psnr_phantom.py (profiles/r01_psnr.json)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))


def test_phantom_psnr_matches_oracle(native_lib):
    import psnr_phantom

    out = psnr_phantom.run("1", n_iter=40, batch=1024, n_samples=32, log=lambda *_: None)
    print(out)
    assert out["abs_diff_inside_db"] <= 0.1 and out["abs_diff_full_db"] <= 0.1
    assert out["rel_l2_ours_vs_oracle_volume"] <= 5e-3
    assert out["psnr_ours_full"] > 10.0  # it did reconstruct something
