"""RigidTransform container and point transforms.

Mirrors the part of nesvor/transform/transform.py that sits on the reconstruction path
(RigidTransform :8-118, mat_first2last/last2first :121-134, ax_* :137-144, *_update_resolution
:147-158, mat_transform_points :259-271, ax_transform_points :274-280, transform_points :283-289).
The Euler / 3-point converters (:161-256) serve SVoRT only and are out of scope.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from .transform_convert import axisangle2mat, mat2axisangle


def _split(mat: torch.Tensor):
    return mat[..., :3], mat[..., 3:]


def mat_first2last(mat: torch.Tensor) -> torch.Tensor:
    R, t = _split(mat)
    return torch.cat([R, torch.matmul(R, t)], -1)


def mat_last2first(mat: torch.Tensor) -> torch.Tensor:
    R, t = _split(mat)
    return torch.cat([R, torch.matmul(R.transpose(-2, -1), t)], -1)


def ax_first2last(axisangle: torch.Tensor) -> torch.Tensor:
    return mat2axisangle(mat_first2last(axisangle2mat(axisangle)).contiguous())


def ax_last2first(axisangle: torch.Tensor) -> torch.Tensor:
    return mat2axisangle(mat_last2first(axisangle2mat(axisangle)).contiguous())


class RigidTransform(object):
    """A batch of rigid transforms stored either as axis-angle rows [n,6] or matrices [n,3,4], in
    the `trans_first` (x -> R(x + t)) or trans-last (x -> Rx + t) convention."""

    def __init__(self, data: torch.Tensor, trans_first: bool = True, device=None) -> None:
        self.trans_first = trans_first
        self._axisangle: Optional[torch.Tensor] = None
        self._matrix: Optional[torch.Tensor] = None
        if device is not None:
            data = data.to(device)
        if data.ndim == 2 and data.shape[1] == 6:
            self._axisangle = data
        elif data.ndim == 3 and data.shape[1] == 3:
            self._matrix = data
        else:
            raise Exception("Unknown format for rigid transform!")

    def _data(self) -> torch.Tensor:
        d = self._axisangle if self._axisangle is not None else self._matrix
        if d is None:
            raise Exception("Both data are None!")
        return d

    def matrix(self, trans_first: bool = True) -> torch.Tensor:
        mat = self._matrix if self._matrix is not None else axisangle2mat(self._axisangle.contiguous())
        if self.trans_first and not trans_first:
            mat = mat_first2last(mat)
        elif not self.trans_first and trans_first:
            mat = mat_last2first(mat)
        return mat

    def axisangle(self, trans_first: bool = True) -> torch.Tensor:
        ax = self._axisangle if self._axisangle is not None else mat2axisangle(self._matrix.contiguous())
        if self.trans_first and not trans_first:
            ax = ax_first2last(ax.contiguous())
        elif not self.trans_first and trans_first:
            ax = ax_last2first(ax.contiguous())
        return ax

    def inv(self) -> "RigidTransform":
        R, t = _split(self.matrix(trans_first=True))
        return RigidTransform(torch.cat((R.transpose(-2, -1), -torch.matmul(R, t)), -1), trans_first=True)

    def compose(self, other: "RigidTransform") -> "RigidTransform":
        R1, t1 = _split(self.matrix(trans_first=True))
        R2, t2 = _split(other.matrix(trans_first=True))
        R = torch.matmul(R1, R2)
        t = t2 + torch.matmul(R2.transpose(-2, -1), t1)
        return RigidTransform(torch.cat((R, t), -1), trans_first=True)

    def __getitem__(self, idx) -> "RigidTransform":
        src = self._data()
        data = src[idx]
        if data.ndim < src.ndim:
            data = data.unsqueeze(0)
        return RigidTransform(data, self.trans_first)

    def detach(self) -> "RigidTransform":
        return RigidTransform(self._data().detach(), self.trans_first)

    def clone(self) -> "RigidTransform":
        return RigidTransform(self._data().clone(), self.trans_first)

    @property
    def device(self):
        return self._data().device

    @staticmethod
    def cat(transforms: Iterable["RigidTransform"]) -> "RigidTransform":
        return RigidTransform(torch.cat([t.matrix(trans_first=True) for t in transforms], 0), trans_first=True)

    def __len__(self) -> int:
        return self._data().shape[0]


def mat_update_resolution(mat: torch.Tensor, res_from, res_to) -> torch.Tensor:
    assert mat.dim() == 3
    fac = torch.ones_like(mat[:1, :1])
    fac[..., 3] = res_from / res_to
    return mat * fac


def ax_update_resolution(ax: torch.Tensor, res_from, res_to) -> torch.Tensor:
    assert ax.dim() == 2
    fac = torch.ones_like(ax[:1])
    fac[:, 3:] = res_from / res_to
    return ax * fac


def mat_transform_points(mat: torch.Tensor, x: torch.Tensor, trans_first: bool) -> torch.Tensor:
    """mat (*,3,4), x (*,3) -> (*,3): R(x + t) if trans_first else Rx + t."""
    R, T = mat[..., :-1], mat[..., -1:]
    x = x[..., None]
    x = torch.matmul(R, x + T) if trans_first else torch.matmul(R, x) + T
    return x[..., 0]


def ax_transform_points(ax: torch.Tensor, x: torch.Tensor, trans_first: bool) -> torch.Tensor:
    mat = axisangle2mat(ax.reshape(-1, 6).contiguous()).view(ax.shape[:-1] + (3, 4))
    return mat_transform_points(mat, x, trans_first)


def transform_points(transform: RigidTransform, x: torch.Tensor) -> torch.Tensor:
    assert x.ndim == 2 and x.shape[-1] == 3
    trans_first = transform.trans_first
    return mat_transform_points(transform.matrix(trans_first), x, trans_first)
