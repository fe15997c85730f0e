"""oracle/inr_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by ``nesvor_b200``).

Pure-PyTorch restatement of the INR half of NeSVoR's hot path, differentiable by autograd, runnable
on CPU in fp32 or fp64.  **Parity unpinned for the hash grid + fp16 MLP**: those live in
tiny-cuda-nn (NVlabs/tiny-cuda-nn, unpinned pip-from-git dependency, /root/reference/README.md:88,
not vendored, not installable here).  This file restates tcnn's published algorithm
(Instant-NGP, Mueller et al. 2022, eqs. 2-4; details in SURVEY.md App. A) and is anchored on the
reference's call sites:

  build_encoding / build_network      /root/reference/nesvor/nesvor/models.py:22-69
  INR.forward                         models.py:142-152
  INR.sample_batch                    models.py:154-174
  NeSVoR.forward / net_forward        models.py:260-355
  NeSVoR.trans_loss                   models.py:357-363
  tv_reg / edge_reg / l2_reg          models.py:366-384
  ax_transform_points & friends       /root/reference/nesvor/transform/transform.py:259-280
  axisangle2mat / mat2axisangle       /root/reference/nesvor/transform/transform_convert_cuda_kernel.cu:15-264
  resolution2sigma                    /root/reference/nesvor/utils/psf.py:5-35
  optimiser / schedule                /root/reference/nesvor/nesvor/train.py:134-165

Rounding model (``emulate_fp16=True``): the CUDA fast path reads an fp16 copy of the table and of
the MLP weights, rounds the encoding output and every hidden activation to fp16 and accumulates in
fp32; last-layer outputs stay fp32 (they are rounded only where they feed the next MLP).  The
oracle reproduces exactly those rounding points so that the 1e-4 rel-L2 parity target on the
rendered pixel ``v_out`` is meaningful; with ``emulate_fp16=False`` it is plain fp32 (or fp64).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

GAUSSIAN_FWHM = 1 / (2 * math.sqrt(2 * math.log(2)))
SINC_FWHM = 1.206709128803223 * GAUSSIAN_FWHM
PRIMES = (1, 2654435761, 805459861)
_U32 = 0xFFFFFFFF


# --------------------------------------------------------------------------------------- hash grid
@dataclass
class GridMeta:
    n_levels: int
    n_features: int
    scale: np.ndarray  # float32 [L]
    res: np.ndarray  # uint32 [L]
    size: np.ndarray  # entries per level [L]
    offset: np.ndarray  # entry offset per level [L+1]
    hashed: np.ndarray  # bool [L]

    @property
    def n_params(self) -> int:
        return int(self.offset[-1]) * self.n_features


def grid_meta(n_levels: int, n_features: int, log2_hashmap_size: int, base_resolution: int, per_level_scale: float) -> GridMeta:
    """Level geometry as tcnn lays it out: scale_l = fp32(base * s^l - 1),
    res_l = ceil(scale_l) + 1, T_l = min(round_up(res_l^3, 8), 2^log2_T)."""
    scale = np.zeros(n_levels, np.float32)
    res = np.zeros(n_levels, np.int64)
    size = np.zeros(n_levels, np.int64)
    log2s = math.log2(float(np.float32(per_level_scale)))  # double precision, narrowed once (platform independent)
    for l in range(n_levels):
        scale[l] = np.float32(2.0 ** (l * log2s) * base_resolution - 1.0)
        res[l] = int(np.ceil(scale[l])) + 1
        dense = min(int(res[l]) ** 3, (2**32 - 1) // 2)
        dense = (dense + 7) // 8 * 8
        size[l] = min(dense, 1 << log2_hashmap_size)
    offset = np.concatenate([[0], np.cumsum(size)])
    hashed = res.astype(object) ** 3 > size
    return GridMeta(n_levels, n_features, scale, res, size, offset, np.asarray(hashed, bool))


def _q16(x: torch.Tensor, on: bool) -> torch.Tensor:
    """Round to fp16 and back (straight-through for autograd) when ``on``."""
    if not on:
        return x
    return x + (x.detach().to(torch.float16).to(x.dtype) - x.detach())


def hashgrid_encode(x: torch.Tensor, table: torch.Tensor, meta: GridMeta, emulate_fp16: bool = False) -> torch.Tensor:
    """x [N,3] in [0,1] (not clamped) -> [N, L*F], level-major.  ``table`` is the flat parameter
    (concat over levels of [T_l, F]).  Differentiable w.r.t. ``table`` and ``x`` (autograd gives
    d enc / d x = scale_l * sum over the other two dims' weights * (feat[+1] - feat[0]))."""
    N = x.shape[0]
    Fe = meta.n_features
    tab = _q16(table, emulate_fp16).view(-1, Fe)
    outs = []
    for l in range(meta.n_levels):
        s = float(meta.scale[l])
        res = int(meta.res[l])
        T = int(meta.size[l])
        # fmaf(scale, x, 0.5): exact product in fp64, one rounding back
        if x.dtype == torch.float32:
            pos = (x.double() * s + 0.5).to(torch.float32)
        else:
            pos = x * s + 0.5
        g = torch.floor(pos.detach())
        w = pos - g
        gi = g.to(torch.int64) & _U32  # (uint32)(int)floorf: negatives wrap
        feat = torch.zeros(N, Fe, dtype=x.dtype, device=x.device)
        for c in range(8):
            bits = [(c >> d) & 1 for d in range(3)]
            cx, cy, cz = [(gi[:, d] + bits[d]) & _U32 for d in range(3)]
            if meta.hashed[l]:
                idx = (cx * PRIMES[0]) ^ ((cy * PRIMES[1]) & _U32) ^ ((cz * PRIMES[2]) & _U32)
                idx = idx & _U32
            else:
                idx = (cx + ((cy * res) & _U32) + ((cz * ((res * res) & _U32)) & _U32)) & _U32
            idx = idx % T + int(meta.offset[l])
            wt = torch.ones(N, dtype=x.dtype, device=x.device)
            for d in range(3):
                wt = wt * (w[:, d] if bits[d] else (1 - w[:, d]))
            feat = feat + wt[:, None] * tab[idx]
        outs.append(feat)
    return _q16(torch.cat(outs, -1), emulate_fp16)


# --------------------------------------------------------------------------------------------- MLP
def mlp_forward(x: torch.Tensor, weights: List[torch.Tensor], biases: Optional[List[Optional[torch.Tensor]]] = None,
                emulate_fp16: bool = False) -> torch.Tensor:
    """ReLU hidden layers, linear output (models.py:28-69).  ``weights[i]`` is [out_i, in_i]; the
    tcnn branch has no biases, the fp32 nn.Sequential branch does.  ``x`` narrower than in_0 is
    zero-padded (tcnn pads to 16)."""
    h = x
    if h.shape[-1] < weights[0].shape[1]:
        h = F.pad(h, (0, weights[0].shape[1] - h.shape[-1]))
    h = _q16(h, emulate_fp16)
    for i, W in enumerate(weights):
        h = h @ _q16(W, emulate_fp16).t()
        if biases is not None and biases[i] is not None:
            h = h + biases[i]
        if i + 1 < len(weights):
            h = _q16(torch.relu(h), emulate_fp16)
    return h


# ------------------------------------------------------------------------------------ rigid poses
def axisangle2mat(ax: torch.Tensor) -> torch.Tensor:
    """[n,6] -> [n,3,4]; Rodrigues with the reference's small-angle branch (theta^2 <= 1e-6 -> I + [w]x)."""
    w = ax[:, :3]
    theta2 = (w * w).sum(-1)
    small = theta2 <= 1e-6
    theta = torch.sqrt(torch.where(small, torch.ones_like(theta2), theta2))
    u = w / theta[:, None]
    s, c = torch.sin(theta), torch.cos(theta)
    oc = 1 - c
    ux, uy, uz = u[:, 0], u[:, 1], u[:, 2]
    R = torch.stack([
        c + ux * ux * oc, ux * uy * oc - uz * s, uy * s + ux * uz * oc,
        uz * s + ux * uy * oc, c + uy * uy * oc, -ux * s + uy * uz * oc,
        -uy * s + ux * uz * oc, ux * s + uy * uz * oc, c + uz * uz * oc], -1).view(-1, 3, 3)
    one, zero = torch.ones_like(theta2), torch.zeros_like(theta2)
    wx, wy, wz = w[:, 0], w[:, 1], w[:, 2]
    Rs = torch.stack([one, -wz, wy, wz, one, -wx, -wy, wx, one], -1).view(-1, 3, 3)
    R = torch.where(small[:, None, None], Rs, R)
    return torch.cat([R, ax[:, 3:, None]], -1)


def mat2axisangle(mat: torch.Tensor) -> torch.Tensor:
    """[n,3,4] -> [n,6] with the reference's four quaternion branches (transform_convert_cuda_kernel.cu:217-251)."""
    R = mat[:, :, :3]
    r = lambda i, j: R[:, i, j]
    d2 = r(2, 2) < 1e-6
    d0_gt_d1 = r(0, 0) > r(1, 1)
    d0_lt_nd1 = r(0, 0) < -r(1, 1)
    safe = lambda t: torch.sqrt(torch.clamp(t, min=1e-30))

    def branch(p):
        if p < 0:
            s = 2 * safe(r(0, 0) + r(1, 1) + r(2, 2) + 1)
            return torch.stack([0.25 * s, (r(2, 1) - r(1, 2)) / s, (r(0, 2) - r(2, 0)) / s, (r(1, 0) - r(0, 1)) / s], -1)
        a, b = (p + 1) % 3, (p + 2) % 3
        o = [i for i in range(3) if i != p]
        s = 2 * safe(r(p, p) - r(o[0], o[0]) - r(o[1], o[1]) + 1)
        comp = [None] * 4
        comp[0] = (r(b, a) - r(a, b)) / s
        comp[1 + p] = 0.25 * s
        comp[1 + a] = (r(p, a) + r(a, p)) / s
        comp[1 + b] = (r(p, b) + r(b, p)) / s
        return torch.stack(comp, -1)

    sel_w = (~d2) & (~d0_lt_nd1)
    sel_x = d2 & d0_gt_d1
    sel_y = d2 & (~d0_gt_d1)
    q = branch(2)
    q = torch.where(sel_y[:, None], branch(1), q)
    q = torch.where(sel_x[:, None], branch(0), q)
    q = torch.where(sel_w[:, None], branch(-1), q)
    q = torch.where((q[:, :1] < 0), -q, q)
    wq, v = q[:, 0], q[:, 1:]
    n2 = (v * v).sum(-1)
    big = n2 > 1e-6
    si = torch.sqrt(torch.where(big, n2, torch.ones_like(n2)))
    theta = 2 * torch.atan2(si, wq)
    fac = torch.where(big, theta / si, 2.0 / wq)
    return torch.cat([v * fac[:, None], mat[:, :, 3]], -1)


def mat_transform_points(mat: torch.Tensor, x: torch.Tensor, trans_first: bool) -> torch.Tensor:
    R, T = mat[..., :-1], mat[..., -1:]
    x = x[..., None]
    x = torch.matmul(R, x + T) if trans_first else torch.matmul(R, x) + T
    return x[..., 0]


def mat_inv(mat: torch.Tensor) -> torch.Tensor:  # RigidTransform.inv, transform.py:46-51 (trans_first matrices)
    R, t = mat[:, :, :3], mat[:, :, 3:]
    return torch.cat((R.transpose(-2, -1), -torch.matmul(R, t)), -1)


def mat_compose(m1: torch.Tensor, m2: torch.Tensor) -> torch.Tensor:  # RigidTransform.compose, transform.py:53-63
    R1, t1, R2, t2 = m1[:, :, :3], m1[:, :, 3:], m2[:, :, :3], m2[:, :, 3:]
    return torch.cat((torch.matmul(R1, R2), t2 + torch.matmul(R2.transpose(-2, -1), t1)), -1)


def resolution2sigma(res: torch.Tensor, isotropic: bool = False) -> torch.Tensor:
    if isotropic:
        return res * GAUSSIAN_FWHM
    return res * torch.tensor([SINC_FWHM, SINC_FWHM, GAUSSIAN_FWHM], dtype=res.dtype, device=res.device)


# ----------------------------------------------------------------------------------------- model
@dataclass
class INRConfig:
    n_levels: int = 12
    n_features_per_level: int = 2
    log2_hashmap_size: int = 19
    base_resolution: int = 7
    level_scale: float = 1.3819
    width: int = 64
    depth: int = 1
    n_features_z: int = 15
    n_features_slice: int = 16
    n_levels_bias: int = 0
    no_transformation_optimization: bool = False
    no_slice_scale: bool = False
    no_pixel_variance: bool = False
    no_slice_variance: bool = False
    image_regularization: str = "edge"
    n_samples: int = 256
    delta: float = 0.2  # already multiplied by v_mean (models.py:192)
    weight_transformation: float = 0.1
    weight_bias: float = 100.0
    weight_image: float = 2.0
    mlp_bias: bool = False  # True for the fp32 nn.Linear branch (models.py:42-67)
    emulate_fp16: bool = False


def _pad16(n: int) -> int:
    return (n + 15) // 16 * 16


def mlp_shapes(n_in: int, n_out: int, width: int, depth: int, pad: bool = True):
    dims = [n_in] + [width] * depth + [n_out]
    if pad:
        dims = [_pad16(d) for d in dims]
    return [(dims[i + 1], dims[i]) for i in range(len(dims) - 1)]


class OracleNeSVoR:
    """Holds the trainable tensors as leaves and evaluates NeSVoR.forward (models.py:260-327)."""

    def __init__(self, cfg: INRConfig, n_slices: int, axisangle: torch.Tensor, resolution: torch.Tensor,
                 bounding_box: torch.Tensor, seed: int = 1337, dtype=torch.float32):
        self.cfg, self.n_slices, self.dtype = cfg, n_slices, dtype
        g = torch.Generator().manual_seed(seed)
        self.meta = grid_meta(cfg.n_levels, cfg.n_features_per_level, cfg.log2_hashmap_size, cfg.base_resolution, cfg.level_scale)
        P: Dict[str, torch.Tensor] = {}
        P["table"] = (torch.rand(self.meta.n_params, generator=g, dtype=torch.float64) * 2e-4 - 1e-4).to(dtype)
        pad = not cfg.mlp_bias
        LF = cfg.n_levels * cfg.n_features_per_level

        def make_mlp(prefix, n_in, n_out):
            for i, (o, k) in enumerate(mlp_shapes(n_in, n_out, cfg.width, cfg.depth, pad)):
                bound = math.sqrt(6.0 / (o + k))
                P[f"{prefix}.w{i}"] = ((torch.rand(o, k, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
                if cfg.mlp_bias:
                    P[f"{prefix}.b{i}"] = ((torch.rand(o, generator=g, dtype=torch.float64) * 2 - 1) / math.sqrt(k)).to(dtype)

        make_mlp("density_net", LF, 1 + cfg.n_features_z)
        if not cfg.no_pixel_variance:
            make_mlp("sigma_net", cfg.n_features_slice + cfg.n_features_z, 1)
        if cfg.n_levels_bias:
            make_mlp("b_net", cfg.n_levels_bias * cfg.n_features_per_level + cfg.n_features_slice, 1)
        if cfg.n_features_slice:
            P["slice_embedding"] = torch.randn(n_slices, cfg.n_features_slice, generator=g, dtype=torch.float64).to(dtype)
        if not cfg.no_slice_scale:
            P["logit_coef"] = torch.zeros(n_slices, dtype=dtype)
        if not cfg.no_slice_variance:
            P["log_var_slice"] = torch.zeros(n_slices, dtype=dtype)
        P["axisangle"] = axisangle.detach().clone().to(dtype)
        self.P = P
        self.axisangle_init = axisangle.detach().clone().to(dtype)
        self.psf_sigma = resolution2sigma(resolution.to(dtype), isotropic=False)
        self.bounding_box = bounding_box.to(dtype)
        self.trainable = [k for k in P if not (k == "axisangle" and cfg.no_transformation_optimization)]
        for k in self.trainable:
            P[k].requires_grad_(True)

    # ------------------------------------------------------------------ pieces
    def _mlp(self, prefix: str, x: torch.Tensor) -> torch.Tensor:
        ws, bs, i = [], [], 0
        while f"{prefix}.w{i}" in self.P:
            ws.append(self.P[f"{prefix}.w{i}"])
            bs.append(self.P.get(f"{prefix}.b{i}"))
            i += 1
        return mlp_forward(x, ws, bs if self.cfg.mlp_bias else None, self.cfg.emulate_fp16)

    def inr_forward(self, x: torch.Tensor):
        """INR.forward (models.py:142-152): world coords [...,3] -> density [...], pe, z."""
        bb = self.bounding_box
        xn = (x - bb[0]) / (bb[1] - bb[0])
        prefix = xn.shape[:-1]
        pe = hashgrid_encode(xn.reshape(-1, 3), self.P["table"], self.meta, self.cfg.emulate_fp16)
        z = self._mlp("density_net", pe)
        density = F.softplus(z[..., 0].view(prefix))
        return density, pe, z

    def trans_loss(self) -> torch.Tensor:
        x = axisangle2mat(self.P["axisangle"])
        y = axisangle2mat(self.axisangle_init)
        err = mat2axisangle(mat_compose(mat_inv(y), x))
        return torch.mean(err[:, :3] ** 2) + 1e-3 * torch.mean(err[:, 3:] ** 2)

    # ------------------------------------------------------------------ NeSVoR.forward
    def forward(self, xyz: torch.Tensor, v: torch.Tensor, slice_idx: torch.Tensor, noise: torch.Tensor,
                return_aux: bool = False):
        cfg, P = self.cfg, self.P
        B, S = xyz.shape[0], noise.shape[1]
        psf_sigma = self.psf_sigma[slice_idx][:, None]
        t = P["axisangle"][slice_idx][:, None]
        pts = xyz[:, None] + noise * psf_sigma
        mat = axisangle2mat(t.reshape(-1, 6)).view(B, 1, 3, 4)
        x = mat_transform_points(mat, pts, True)
        se = P["slice_embedding"][slice_idx][:, None].expand(-1, S, -1) if cfg.n_features_slice else None
        density, pe, z = self.inr_forward(x)
        zs = [se.reshape(-1, se.shape[-1])] if se is not None else []
        log_bias = log_var = None
        nz = 1 + cfg.n_features_z
        if cfg.n_levels_bias:
            pe_bias = pe[..., : cfg.n_levels_bias * cfg.n_features_per_level]
            log_bias = self._mlp("b_net", torch.cat(zs + [pe_bias], -1))[..., 0].view(B, S)
        if not cfg.no_pixel_variance:
            log_var = self._mlp("sigma_net", torch.cat(zs + [z[..., 1:nz]], -1))[..., 0].view(B, S)
        bias = log_bias.exp() if log_bias is not None else 1
        bias_detach = bias.detach() if log_bias is not None else 1
        var = log_var.exp() if log_var is not None else 1
        c = F.softmax(P["logit_coef"], 0)[slice_idx] * self.n_slices if not cfg.no_slice_scale else 1
        v_out = c * (bias * density).mean(-1)
        if not cfg.no_pixel_variance:
            var = (bias_detach * var).mean(-1)
            var = (c.detach() if not cfg.no_slice_scale else 1) * var
            var = var**2
        if not cfg.no_slice_variance:
            var = var + P["log_var_slice"].exp()[slice_idx]
        losses = {"MSE": ((v_out - v) ** 2 / (2 * var)).mean()}
        if not (cfg.no_pixel_variance and cfg.no_slice_variance):
            losses["logVar"] = 0.5 * var.log().mean()
            losses["MSE+logVar"] = losses["MSE"] + losses["logVar"]
        if not cfg.no_transformation_optimization:
            losses["transReg"] = self.trans_loss()
        if cfg.n_levels_bias:
            losses["biasReg"] = log_bias.mean() ** 2
        losses["imageReg"] = image_reg(cfg.image_regularization, density, x, cfg.delta)
        if return_aux:
            return losses, {"v_out": v_out, "density": density, "x": x, "var": var}
        return losses

    def total_loss(self, losses: Dict[str, torch.Tensor]) -> torch.Tensor:
        cfg = self.cfg
        wts = {"MSE": 1.0, "logVar": 1.0, "transReg": cfg.weight_transformation, "biasReg": cfg.weight_bias,
               "imageReg": cfg.weight_image}
        total = 0
        for k, val in losses.items():
            if k in wts and wts[k]:
                total = total + wts[k] * val
        return total

    # ------------------------------------------------------------------ inference (sample.py:17-53)
    def render(self, xyz: torch.Tensor, noise: Optional[torch.Tensor], psf_sigma, mat: Optional[torch.Tensor] = None):
        """INR.sample_batch + INR.forward(...).mean(-1): xyz [M,3], noise [M,S,3] or None."""
        if noise is not None:
            sig = psf_sigma.view(-1, 1, 3) if isinstance(psf_sigma, torch.Tensor) and psf_sigma.ndim > 0 else psf_sigma
            pts = xyz[:, None] + noise * sig
        else:
            pts = xyz[:, None]
        if mat is not None:
            pts = mat_transform_points(mat[:, None], pts, True)
        density, _, _ = self.inr_forward(pts)
        return density.mean(-1)


def image_reg(kind: str, density: torch.Tensor, xyz: torch.Tensor, delta: float) -> torch.Tensor:
    d_density = density - torch.flip(density, (1,))
    dx2 = ((xyz - torch.flip(xyz, (1,))) ** 2).sum(-1) + 1e-6
    if kind == "TV":
        return torch.abs(d_density / dx2.sqrt()).mean()
    if kind == "edge":
        return delta * ((1 + d_density**2 / dx2 / (delta * delta)).sqrt().mean() - 1)
    if kind == "L2":
        return (d_density**2 / dx2).mean()
    raise ValueError(kind)


def make_optimizer(model: OracleNeSVoR, lr: float = 5e-3):
    """train.py:134-152: AdamW, betas (0.9, 0.99), eps 1e-15, two groups that both end up at wd 1e-2."""
    net = [model.P[k] for k in model.trainable if "_net" in k]
    enc = [model.P[k] for k in model.trainable if "_net" not in k]
    return torch.optim.AdamW([{"params": enc}, {"params": net, "weight_decay": 1e-2}], lr=lr, betas=(0.9, 0.99), eps=1e-15)
