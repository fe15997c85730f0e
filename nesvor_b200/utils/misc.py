"""Small helpers used by the drivers of the hot path (semantics of nesvor/utils/misc.py:29-122)."""
import collections.abc
from typing import Any, Dict, Optional

import torch
import torch.nn.functional as F


def meshgrid(shape_xyz, resolution_xyz, min_xyz=None, device=None, stack_output: bool = True):
    """Centred voxel-centre coordinates, (x, y, z) last, array order (z, y, x) (misc.py:29-60)."""
    assert len(shape_xyz) == len(resolution_xyz)
    if min_xyz is None:
        min_xyz = tuple(-(s - 1) * r / 2 for s, r in zip(shape_xyz, resolution_xyz))
    if device is None:
        if isinstance(shape_xyz, torch.Tensor):
            device = shape_xyz.device
        elif isinstance(resolution_xyz, torch.Tensor):
            device = resolution_xyz.device
        else:
            device = torch.device("cpu")
    arr = [torch.arange(int(s), dtype=torch.float32, device=device) * r + m for s, r, m in zip(shape_xyz, resolution_xyz, min_xyz)]
    grid = torch.meshgrid(arr[::-1], indexing="ij")[::-1]
    return torch.stack(grid, -1) if stack_output else grid


def gaussian_1d_kernel(sigma: float, truncated: float, device) -> torch.Tensor:
    tail = int(max(sigma * truncated, 0.5) + 0.5)
    x = torch.arange(-tail, tail + 1, dtype=torch.float, device=device)
    t = 0.70710678 / sigma
    return (0.5 * ((t * (x + 0.5)).erf() - (t * (x - 0.5)).erf())).clamp(min=0)


def gaussian_blur(x: torch.Tensor, sigma, truncated: float) -> torch.Tensor:
    """Separable Gaussian blur of [N,C,*spatial] (misc.py:63-88)."""
    nd = x.ndim - 2
    if not isinstance(sigma, collections.abc.Iterable):
        sigma = [sigma] * nd
    conv = [F.conv1d, F.conv2d, F.conv3d][nd - 1]
    c = x.shape[1]
    for d, s in enumerate(sigma):
        k = gaussian_1d_kernel(float(s), truncated, x.device)
        shape = [1] * x.ndim
        shape[d + 2] = -1
        k = k.reshape(shape).repeat(*([c, 1] + [1] * nd))
        pad = [0] * nd
        pad[d] = (k.shape[d + 2] - 1) // 2
        x = conv(x, k, padding=pad, groups=c)
    return x


class MovingAverage:
    """Bias-corrected exponential moving average per key (misc.py:91-122)."""

    def __init__(self, alpha: float) -> None:
        assert 0 <= alpha < 1
        self.alpha = alpha
        self._value: Dict[str, Any] = {}

    def __call__(self, key: str, value) -> None:
        num, total = self._value.get(key, (0, 0.0))
        self._value[key] = (num + 1, total * self.alpha + (1 - self.alpha) * value)

    def __getitem__(self, key: str):
        if key not in self._value:
            return 0
        num, v = self._value[key]
        return v / (1 - self.alpha**num) if self.alpha else v

    def __str__(self) -> str:
        return ", ".join("%s = %.3e" % (k, self[k]) for k in self._value)
