"""Native replacements for the two tiny-cuda-nn modules NeSVoR instantiates.

* ``HashGridEncoding``  <- tcnn.Encoding(n_input_dims, {"otype": "HashGrid", ...}, dtype)
  (nesvor/nesvor/models.py:22-25, built :102-111, called :146)
* ``FusedMLP``          <- tcnn.Network(n_input_dims, n_output_dims, {"otype": "CutlassMLP", ...})
  (nesvor/nesvor/models.py:30-41)

Both keep tcnn's module surface: a single flat fp32 parameter ``params`` (state-dict key
``params``), fp16 compute when ``dtype == torch.float16``, an internal loss scale of 128 on the
backward operands, and ``n_output_dims`` / ``n_input_dims`` attributes.  They are differentiable
w.r.t. ``params`` and their input.  tiny-cuda-nn is not part of /root/reference; the semantics are
those documented in SURVEY.md App. A.
"""
import ctypes
import math

import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _lib

LOSS_SCALE = 128.0


def _pad(n: int, m: int = 16) -> int:
    return (n + m - 1) // m * m


class _HashGridFunction(Function):
    @staticmethod
    def forward(ctx, x, params, module):
        _lib.require_cuda("x", x, torch.float32)
        _lib.require_cuda("params", params, torch.float32)
        N = x.shape[0]
        meta = module.meta
        half = module.dtype == torch.float16
        table = params.to(torch.float16) if half else params
        out = torch.empty((N, module.n_output_dims), dtype=module.dtype, device=x.device)
        with torch.cuda.device(x.device):
            fn = _lib.lib().nsv_hashgrid_fwd_f16 if half else _lib.lib().nsv_hashgrid_fwd_f32
            rc = fn(_lib.ptr(x), _lib.ptr(table), ctypes.byref(meta), _lib.ptr(out), ctypes.c_int64(N), _lib.stream(x.device))
        _lib.check(rc, "nsv_hashgrid_fwd")
        ctx.save_for_backward(x, params)
        ctx.module = module
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, params = ctx.saved_tensors
        module = ctx.module
        meta = module.meta
        N = x.shape[0]
        half = module.dtype == torch.float16
        lib = _lib.lib()
        grad_x = grad_params = None
        if half:  # tcnn scales the incoming fp16 gradient by its loss scale and unscales the results
            go = (grad_out * LOSS_SCALE).to(torch.float16).contiguous()
        else:
            go = grad_out.to(torch.float32).contiguous()
        with torch.cuda.device(x.device):
            st = _lib.stream(x.device)
            if ctx.needs_input_grad[1]:
                grad_params = torch.zeros_like(params)
                if half:
                    rc = lib.nsv_hashgrid_bwd_params_f16(_lib.ptr(x), _lib.ptr(go), ctypes.byref(meta), _lib.ptr(grad_params),
                                                         ctypes.c_float(LOSS_SCALE), ctypes.c_int64(N), st)
                else:
                    rc = lib.nsv_hashgrid_bwd_params_f32(_lib.ptr(x), _lib.ptr(go), ctypes.byref(meta), _lib.ptr(grad_params),
                                                         ctypes.c_int64(N), st)
                _lib.check(rc, "nsv_hashgrid_bwd_params")
            if ctx.needs_input_grad[0]:
                grad_x = torch.empty_like(x)
                if half:
                    table = params.to(torch.float16)
                    rc = lib.nsv_hashgrid_bwd_input_f16(_lib.ptr(x), _lib.ptr(table), _lib.ptr(go), ctypes.byref(meta),
                                                        ctypes.c_float(LOSS_SCALE), _lib.ptr(grad_x), ctypes.c_int64(N), st)
                else:
                    rc = lib.nsv_hashgrid_bwd_input_f32(_lib.ptr(x), _lib.ptr(params), _lib.ptr(go), ctypes.byref(meta),
                                                        _lib.ptr(grad_x), ctypes.c_int64(N), st)
                _lib.check(rc, "nsv_hashgrid_bwd_input")
        return grad_x, grad_params, None


class HashGridEncoding(nn.Module):
    """[N, 3] fp32 in [0, 1] -> [N, n_levels * n_features_per_level] (level-major)."""

    def __init__(self, n_input_dims: int, encoding_config: dict, dtype=torch.float16, seed: int = 1337):
        super().__init__()
        if n_input_dims != 3:
            raise ValueError("HashGridEncoding: only 3-D inputs are supported (NeSVoR encodes xyz)")
        if encoding_config.get("otype", "HashGrid") != "HashGrid":
            raise ValueError(f"unsupported encoding otype {encoding_config.get('otype')}")
        self.n_input_dims = n_input_dims
        self.n_levels = int(encoding_config["n_levels"])
        self.n_features_per_level = int(encoding_config.get("n_features_per_level", 2))
        self.log2_hashmap_size = int(encoding_config.get("log2_hashmap_size", 19))
        self.base_resolution = int(encoding_config.get("base_resolution", 16))
        self.per_level_scale = float(encoding_config.get("per_level_scale", 2.0))
        self.dtype = dtype
        self.meta, n_entries = _lib.make_grid_meta(self.n_levels, self.n_features_per_level, self.log2_hashmap_size,
                                                   self.base_resolution, self.per_level_scale)
        self.n_output_dims = self.n_levels * self.n_features_per_level
        g = torch.Generator().manual_seed(seed)
        init = torch.rand(n_entries * self.n_features_per_level, generator=g, dtype=torch.float32) * 2e-4 - 1e-4
        self.params = nn.Parameter(init)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _HashGridFunction.apply(x.to(torch.float32).contiguous(), self.params, self)

    def extra_repr(self) -> str:
        return (f"n_levels={self.n_levels}, F={self.n_features_per_level}, log2_T={self.log2_hashmap_size}, "
                f"base={self.base_resolution}, scale={self.per_level_scale}, dtype={self.dtype}")


class _FusedMLPFunction(Function):
    @staticmethod
    def forward(ctx, x, params, module):
        N = x.shape[0]
        w16 = params.to(torch.float16)
        x16 = x.to(torch.float16)
        if x16.shape[1] != module.n_in_padded:
            x16 = torch.nn.functional.pad(x16, (0, module.n_in_padded - x16.shape[1]))
        x16 = x16.contiguous()
        _lib.require_cuda("x", x16)
        out = torch.empty((N, module.n_out_padded), dtype=torch.float16, device=x.device)
        hidden = torch.empty((module.n_hidden_layers, N, module.n_neurons), dtype=torch.float16, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().nsv_mlp_fwd_f16(_lib.ptr(x16), _lib.ptr(w16), _lib.ptr(out), _lib.ptr(hidden), ctypes.c_int64(N),
                                            ctypes.c_int(module.n_in_padded), ctypes.c_int(module.n_out_padded),
                                            ctypes.c_int(module.n_neurons), ctypes.c_int(module.n_hidden_layers),
                                            _lib.stream(x.device))
        _lib.check(rc, "nsv_mlp_fwd_f16")
        ctx.save_for_backward(x16, params, hidden)
        ctx.module = module
        ctx.in_dtype = x.dtype
        ctx.n_in = x.shape[1]
        return out[:, : module.n_output_dims]

    @staticmethod
    def backward(ctx, grad_out):
        x16, params, hidden = ctx.saved_tensors
        module = ctx.module
        N = x16.shape[0]
        go = torch.zeros((N, module.n_out_padded), dtype=torch.float16, device=x16.device)
        go[:, : module.n_output_dims] = (grad_out.float() * LOSS_SCALE).to(torch.float16)
        w16 = params.to(torch.float16)
        grad_w = torch.zeros_like(params)
        grad_x = torch.empty_like(x16) if ctx.needs_input_grad[0] else None
        with torch.cuda.device(x16.device):
            rc = _lib.lib().nsv_mlp_bwd_f16(_lib.ptr(x16), _lib.ptr(w16), _lib.ptr(hidden), _lib.ptr(go), _lib.ptr(grad_x),
                                            _lib.ptr(grad_w), ctypes.c_int64(N), ctypes.c_int(module.n_in_padded),
                                            ctypes.c_int(module.n_out_padded), ctypes.c_int(module.n_neurons),
                                            ctypes.c_int(module.n_hidden_layers), _lib.stream(x16.device))
        _lib.check(rc, "nsv_mlp_bwd_f16")
        grad_w = grad_w / LOSS_SCALE
        if grad_x is not None:
            grad_x = (grad_x[:, : ctx.n_in].float() / LOSS_SCALE).to(ctx.in_dtype)
        return grad_x, grad_w, None


class FusedMLP(nn.Module):
    """ReLU MLP without biases on fp16 tensor cores; returns fp16 like tcnn.Network."""

    def __init__(self, n_input_dims: int, n_output_dims: int, network_config: dict, seed: int = 1337):
        super().__init__()
        if network_config.get("activation", "ReLU") != "ReLU" or network_config.get("output_activation", "None") != "None":
            raise ValueError("FusedMLP supports activation='ReLU', output_activation='None' (all NeSVoR uses)")
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims
        self.n_neurons = int(network_config["n_neurons"])
        self.n_hidden_layers = int(network_config["n_hidden_layers"])
        if self.n_hidden_layers < 1:
            raise ValueError("FusedMLP needs at least one hidden layer")
        self.n_in_padded = 32 if n_input_dims <= 32 else 64  # kernel instantiations: 32 | 64 input columns
        if n_input_dims > 64:
            raise ValueError("FusedMLP: n_input_dims > 64 is not instantiated")
        self.n_out_padded = _pad(n_output_dims)
        dims = [self.n_in_padded] + [self.n_neurons] * self.n_hidden_layers + [self.n_out_padded]
        self.layer_shapes = [(dims[i + 1], dims[i]) for i in range(len(dims) - 1)]
        g = torch.Generator().manual_seed(seed)
        logical_in = [_pad(n_input_dims)] + [self.n_neurons] * self.n_hidden_layers
        chunks = []
        for (o, k), kin in zip(self.layer_shapes, logical_in):
            bound = math.sqrt(6.0 / (o + kin))  # Xavier uniform on the (16-padded) logical shape
            w = (torch.rand(o, k, generator=g, dtype=torch.float32) * 2 - 1) * bound
            w[:, kin:] = 0  # columns beyond tcnn's 16-padded width only ever meet zero inputs: kept at zero (gradient 0, decay of 0)
            chunks.append(w.reshape(-1))
        self.params = nn.Parameter(torch.cat(chunks))

    def weight_views(self):
        """Per-layer [out, in] views into the flat parameter (input columns beyond n_input_dims multiply zeros)."""
        views, off = [], 0
        for o, k in self.layer_shapes:
            views.append(self.params[off : off + o * k].view(o, k))
            off += o * k
        return views

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _FusedMLPFunction.apply(x, self.params, self)
