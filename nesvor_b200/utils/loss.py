"""Normalised cross-correlation between simulated and acquired slices (nesvor/utils/loss.py:6-69, restated).

`ncc_loss(I, J, mask, win, level, eps, reduction)` returns MINUS the squared correlation coefficient
    cc = cov(I, J)^2 / (var(I) var(J) + eps)
either globally per sample (`win=None`; with a mask the moments are sums over the mask divided by its size + eps) or
locally in a `win`-wide box window (moments by a mean filter, window shrunk by 2^level and forced odd).
"""
from typing import Optional

import torch
import torch.nn.functional as F


def _moments_global(I: torch.Tensor, J: torch.Tensor, mask: Optional[torch.Tensor], eps: float):
    I, J = I.flatten(1), J.flatten(1)
    if mask is None:
        mean = lambda t: t.mean(-1)  # noqa: E731
    else:
        n = mask.flatten(1).sum(-1) + eps
        mean = lambda t: t.sum(-1) / n  # noqa: E731
    return mean(I), mean(J), mean(I * I), mean(J * J), mean(I * J)


def _moments_local(I: torch.Tensor, J: torch.Tensor, win: int, dims: int):
    box = torch.full([1, 1] + [win] * dims, 1.0 / win**dims, device=I.device, dtype=I.dtype)
    conv = (F.conv1d, F.conv2d, F.conv3d)[dims - 1]
    mean = lambda t: conv(t, box, stride=1, padding=win // 2)  # noqa: E731
    return mean(I), mean(J), mean(I * I), mean(J * J), mean(I * J)


def ncc_loss(I: torch.Tensor, J: torch.Tensor, mask: Optional[torch.Tensor] = None, win: Optional[int] = 9, level: int = 0,
             eps: float = 1e-6, reduction: str = "none") -> torch.Tensor:
    dims, channels = I.ndim - 2, I.shape[1]
    if mask is not None:
        I, J = I * mask, J * mask
    if win is None:
        mi, mj, mii, mjj, mij = _moments_global(I, J, mask, eps)
        out_shape = (-1, channels)
    else:
        I, J = I.reshape(-1, 1, *I.shape[2:]), J.reshape(-1, 1, *J.shape[2:])
        win = 2 * int(win / 2**level / 2) + 1
        mi, mj, mii, mjj, mij = _moments_local(I, J, win, dims)
        out_shape = (-1, channels, *I.shape[2:])
    cov = mij - mi * mj
    cc = cov * cov / ((mii - mi * mi) * (mjj - mj * mj) + eps)
    if reduction == "mean":
        return -cc.mean()
    if reduction == "sum":
        return -cc.sum()
    return -cc.view(out_shape)
