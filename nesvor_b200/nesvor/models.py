"""INR / NeSVoR models on the native B200 kernels.

Host-side mirror of nesvor/nesvor/models.py: same names (`build_encoding`, `build_network`, `INR`
(alias `INRModel`), `NeSVoR`, loss keys, `tv_reg` / `edge_reg` / `l2_reg`), same constructor and
`forward` signatures, same state-dict keys (`bounding_box`, `encoding.params`, `density_net.*`).
tiny-cuda-nn is replaced by `HashGridEncoding` / `FusedMLP` (encoding.py); the per-iteration op
sequence of `NeSVoR.forward` (models.py:260-327) is either composed from those modules under
autograd (`fused=False`, every head supported, fp32 or fp16) or executed by the single fused
training kernel `nsv_inr_train_step` (`fused=True`, see fused.py).
"""
from argparse import Namespace
from math import log2
from typing import Any, Dict, Optional, Union
import logging

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..transform import RigidTransform, ax_transform_points, mat_transform_points
from ..utils import resolution2sigma
from .encoding import FusedMLP, HashGridEncoding

# keys for losses / regularisers (models.py:13-19)
D_LOSS = "MSE"
S_LOSS = "logVar"
DS_LOSS = "MSE+logVar"
B_REG = "biasReg"
T_REG = "transReg"
I_REG = "imageReg"


def build_encoding(**config):
    """tcnn.Encoding(n_input_dims, encoding_config, dtype) stand-in (models.py:22-25)."""
    n_input_dims = config.pop("n_input_dims")
    dtype = config.pop("dtype")
    return HashGridEncoding(n_input_dims=n_input_dims, encoding_config=config, dtype=dtype)


def build_network(**config):
    """fp16 -> fused tensor-core MLP without biases (tcnn.Network stand-in, models.py:30-41);
    fp32 -> nn.Sequential of nn.Linear (+bias) / activations, exactly models.py:42-67."""
    dtype = config.pop("dtype")
    if dtype == torch.float16:
        return FusedMLP(
            n_input_dims=config["n_input_dims"],
            n_output_dims=config["n_output_dims"],
            network_config={
                "otype": "CutlassMLP",
                "activation": config["activation"],
                "output_activation": config["output_activation"],
                "n_neurons": config["n_neurons"],
                "n_hidden_layers": config["n_hidden_layers"],
            },
        )
    if dtype == torch.float32:
        act = None if config["activation"] == "None" else getattr(nn, config["activation"])
        out_act = None if config["output_activation"] == "None" else getattr(nn, config["output_activation"])
        layers = []
        n_in, width, depth = config["n_input_dims"], config["n_neurons"], config["n_hidden_layers"]
        if depth > 0:
            layers.append(nn.Linear(n_in, width))
            for _ in range(depth - 1):
                if act is not None:
                    layers.append(act())
                layers.append(nn.Linear(width, width))
            if act is not None:
                layers.append(act())
            layers.append(nn.Linear(width, config["n_output_dims"]))
        else:
            layers.append(nn.Linear(n_in, config["n_output_dims"]))
        if out_act is not None:
            layers.append(out_act())
        return nn.Sequential(*layers)
    raise ValueError("unknown dtype")


def hashgrid_hyperparameters(bounding_box: torch.Tensor, args: Namespace):
    """base resolution and level count from the reconstruction extent (models.py:79-101)."""
    extent = (bounding_box[1] - bounding_box[0]).max()
    base_resolution = int((extent / args.coarsest_resolution).ceil().int().item())
    n_levels = int((torch.log2(extent / args.finest_resolution / base_resolution) / log2(args.level_scale) + 1).ceil().int().item())
    return base_resolution, n_levels


class INR(nn.Module):
    def __init__(self, bounding_box: torch.Tensor, args: Namespace) -> None:
        super().__init__()
        self.register_buffer("bounding_box", bounding_box)
        base_resolution, n_levels = hashgrid_hyperparameters(self.bounding_box, args)
        # explicit overrides (BASELINE config 2 fixes L instead of deriving it from resolutions)
        n_levels = int(getattr(args, "n_levels", None) or n_levels)
        base_resolution = int(getattr(args, "base_resolution", None) or base_resolution)
        self.encoding = build_encoding(
            n_input_dims=3,
            otype="HashGrid",
            n_levels=n_levels,
            n_features_per_level=args.n_features_per_level,
            log2_hashmap_size=args.log2_hashmap_size,
            base_resolution=base_resolution,
            per_level_scale=args.level_scale,
            dtype=args.dtype,
        )
        self.density_net = build_network(
            n_input_dims=n_levels * args.n_features_per_level,
            n_output_dims=1 + args.n_features_z,
            activation="ReLU",
            output_activation="None",
            n_neurons=args.width,
            n_hidden_layers=args.depth,
            dtype=args.dtype,
        )
        logging.debug(
            "hyperparameters for hash grid encoding: lowest_grid_size=%d, highest_grid_size=%d, scale=%1.2f, n_levels=%d",
            base_resolution, int(base_resolution * args.level_scale ** (n_levels - 1)), args.level_scale, n_levels)

    def forward(self, x: torch.Tensor, return_all: bool = True):
        x = (x - self.bounding_box[0]) / (self.bounding_box[1] - self.bounding_box[0])
        prefix_shape = x.shape[:-1]
        x = x.view(-1, x.shape[-1])
        pe = self.encoding(x)
        z = self.density_net(pe)
        density = F.softplus(z[..., 0].view(prefix_shape).float())
        if return_all:
            return density, pe, z
        return density

    def sample_batch(self, xyz: torch.Tensor, transformation: Optional[RigidTransform],
                     psf_sigma: Union[float, torch.Tensor], n_samples: int) -> torch.Tensor:
        if n_samples > 1:
            if isinstance(psf_sigma, torch.Tensor):
                psf_sigma = psf_sigma.view(-1, 1, 3)
            xyz_psf = torch.randn(xyz.shape[0], n_samples, 3, dtype=xyz.dtype, device=xyz.device)
            xyz = xyz[:, None] + xyz_psf * psf_sigma
        else:
            xyz = xyz[:, None]
        if transformation is not None:
            trans_first = transformation.trans_first
            mat = transformation.matrix(trans_first)
            xyz = mat_transform_points(mat[:, None], xyz, trans_first)
        return xyz


INRModel = INR  # the name BASELINE.json's north star uses


class NeSVoR(nn.Module):
    def __init__(self, transformation: RigidTransform, resolution: torch.Tensor, v_mean: float,
                 bounding_box: torch.Tensor, args: Namespace) -> None:
        super().__init__()
        self.args = args
        self.n_slices = 0
        self.trans_first = True
        self.transformation = transformation
        self.psf_sigma = resolution2sigma(resolution, isotropic=False)
        self.delta = args.delta * v_mean
        self.image_regularization = {"TV": tv_reg, "edge": edge_reg, "L2": l2_reg}[args.image_regularization]
        self.build_network(bounding_box)
        self.to(args.device)
        self.psf_sigma = self.psf_sigma.to(args.device)

    @property
    def transformation(self) -> RigidTransform:
        return RigidTransform(self.axisangle.detach(), self.trans_first)

    @transformation.setter
    def transformation(self, value: RigidTransform) -> None:
        if self.n_slices == 0:
            self.n_slices = len(value)
        else:
            assert self.n_slices == len(value)
        axisangle = value.axisangle(self.trans_first)
        self.register_buffer("axisangle_init", axisangle.detach().clone())
        if not self.args.no_transformation_optimization:
            self.axisangle = nn.Parameter(axisangle.detach().clone())
        else:
            self.register_buffer("axisangle", axisangle.detach().clone())

    def build_network(self, bounding_box) -> None:
        a = self.args
        if a.n_features_slice:
            self.slice_embedding = nn.Embedding(self.n_slices, a.n_features_slice)
        if not a.no_slice_scale:
            self.logit_coef = nn.Parameter(torch.zeros(self.n_slices, dtype=torch.float32))
        if not a.no_slice_variance:
            self.log_var_slice = nn.Parameter(torch.zeros(self.n_slices, dtype=torch.float32))
        self.inr = INR(bounding_box, a)
        if not a.no_pixel_variance:
            self.sigma_net = build_network(
                n_input_dims=a.n_features_slice + a.n_features_z, n_output_dims=1, activation="ReLU",
                output_activation="None", n_neurons=a.width, n_hidden_layers=a.depth, dtype=a.dtype)
        if a.n_levels_bias:
            self.b_net = build_network(
                n_input_dims=a.n_levels_bias * a.n_features_per_level + a.n_features_slice, n_output_dims=1,
                activation="ReLU", output_activation="None", n_neurons=a.width, n_hidden_layers=a.depth, dtype=a.dtype)

    def forward(self, xyz: torch.Tensor, v: torch.Tensor, slice_idx: torch.Tensor,
                noise: Optional[torch.Tensor] = None, return_v_out: bool = False) -> Dict[str, Any]:
        """`noise` (B, n_samples, 3) replaces the internal torch.randn draw (parity tests);
        `return_v_out` adds the rendered pixel under key "v_out" (not a loss)."""
        a = self.args
        batch_size, n_samples = xyz.shape[0], a.n_samples
        xyz_psf = noise if noise is not None else torch.randn(batch_size, n_samples, 3, dtype=xyz.dtype, device=xyz.device)
        psf_sigma = self.psf_sigma[slice_idx][:, None]
        t = self.axisangle[slice_idx][:, None]
        xyz = ax_transform_points(t, xyz[:, None] + xyz_psf * psf_sigma, self.trans_first)
        se = self.slice_embedding(slice_idx)[:, None].expand(-1, n_samples, -1) if a.n_features_slice else None
        results = self.net_forward(xyz, se)
        density = results["density"]
        if "log_bias" in results:
            log_bias = results["log_bias"].float()
            bias = log_bias.exp()
            bias_detach = bias.detach()
        else:
            log_bias, bias, bias_detach = 0, 1, 1
        var = results["log_var"].float().exp() if "log_var" in results else 1
        c: Any = F.softmax(self.logit_coef, 0)[slice_idx] * self.n_slices if not a.no_slice_scale else 1
        v_out = c * (bias * density).mean(-1)
        if not a.no_pixel_variance:
            var = (bias_detach * var).mean(-1)
            var = (c.detach() if not a.no_slice_scale else 1) * var
            var = var**2
        if not a.no_slice_variance:
            var = var + self.log_var_slice.exp()[slice_idx]
        losses = {D_LOSS: ((v_out - v) ** 2 / (2 * var)).mean()}
        if not (a.no_pixel_variance and a.no_slice_variance):
            losses[S_LOSS] = 0.5 * var.log().mean()
            losses[DS_LOSS] = losses[D_LOSS] + losses[S_LOSS]
        if not a.no_transformation_optimization:
            losses[T_REG] = self.trans_loss(trans_first=self.trans_first)
        if a.n_levels_bias:
            losses[B_REG] = log_bias.mean() ** 2
        losses[I_REG] = self.image_regularization(density, xyz, self.delta)
        if return_v_out:
            losses["v_out"] = v_out
        return losses

    def net_forward(self, x: torch.Tensor, se: Optional[torch.Tensor] = None) -> Dict[str, Any]:
        a = self.args
        density, pe, z = self.inr(x)
        prefix_shape = density.shape
        results = {"density": density}
        zs = []
        if se is not None:
            zs.append(se.reshape(-1, se.shape[-1]))
        if a.n_levels_bias:
            pe_bias = pe[..., : a.n_levels_bias * a.n_features_per_level]
            results["log_bias"] = self.b_net(torch.cat([t.to(pe.dtype) for t in zs] + [pe_bias], -1)).view(prefix_shape)
        if not a.no_pixel_variance:
            zs.append(z[..., 1:])
            results["log_var"] = self.sigma_net(torch.cat([t.to(z.dtype) for t in zs], -1)).view(prefix_shape)
        return results

    def trans_loss(self, trans_first: bool = True) -> torch.Tensor:
        x = RigidTransform(self.axisangle, trans_first=trans_first)
        y = RigidTransform(self.axisangle_init, trans_first=trans_first)
        err = y.inv().compose(x).axisangle(trans_first=trans_first)
        return torch.mean(err[:, :3] ** 2) + 1e-3 * torch.mean(err[:, 3:] ** 2)


def _pair_terms(density: torch.Tensor, xyz: torch.Tensor):
    d_density = density - torch.flip(density, (1,))
    dx2 = ((xyz - torch.flip(xyz, (1,))) ** 2).sum(-1) + 1e-6
    return d_density, dx2


def tv_reg(density: torch.Tensor, xyz: torch.Tensor, delta: float):
    d_density, dx2 = _pair_terms(density, xyz)
    return torch.abs(d_density / dx2.sqrt()).mean()


def edge_reg(density: torch.Tensor, xyz: torch.Tensor, delta: float):
    d_density, dx2 = _pair_terms(density, xyz)
    return delta * ((1 + d_density**2 / dx2 / (delta * delta)).sqrt().mean() - 1)


def l2_reg(density: torch.Tensor, xyz: torch.Tensor, delta: float):
    d_density, dx2 = _pair_terms(density, xyz)
    return (d_density**2 / dx2).mean()
