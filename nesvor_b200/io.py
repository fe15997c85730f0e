"""Model checkpoints of the INR path: `model.pt` = {"model": INR.state_dict(), "mask": Volume, "args": Namespace}
(nesvor/cli/io.py:36-59).

`save_model` writes exactly that dictionary.  `load_model` also reads checkpoints written by the reference: its
pickled helper objects (`nesvor.image.image.Volume`, `nesvor.transform.transform.RigidTransform`) are mapped onto this
package's classes (same attribute names), and tiny-cuda-nn's flat parameters are taken over as they are -- the hash
table is the same level-major `[sum_l T_l, F]` fp32 vector and the MLP the same row-major `[out, in]` layers; only a
first layer whose input tcnn pads to 16 columns is widened to the 32 columns the kernels are instantiated for.
NIfTI volumes / slice folders (nibabel) stay out of scope.
"""
import pickle
import types
from argparse import Namespace
from typing import Optional, Tuple

import torch

from .image import Volume
from .nesvor.models import INR

_CLASS_MAP = {
    ("nesvor.image.image", "Volume"): ("nesvor_b200.image.image", "Volume"),
    ("nesvor.image.image", "Slice"): ("nesvor_b200.image.image", "Slice"),
    ("nesvor.image.image", "Image"): ("nesvor_b200.image.image", "Image"),
    ("nesvor.image", "Volume"): ("nesvor_b200.image.image", "Volume"),
    ("nesvor.image", "Slice"): ("nesvor_b200.image.image", "Slice"),
    ("nesvor.transform.transform", "RigidTransform"): ("nesvor_b200.transform.transform", "RigidTransform"),
    ("nesvor.transform", "RigidTransform"): ("nesvor_b200.transform.transform", "RigidTransform"),
}


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        module, name = _CLASS_MAP.get((module, name), (module, name))
        return super().find_class(module, name)


_pickle_module = types.ModuleType("nesvor_b200._checkpoint_pickle")
_pickle_module.__dict__.update({k: getattr(pickle, k) for k in dir(pickle) if not k.startswith("__")})
_pickle_module.Unpickler = _Unpickler
_pickle_module.load = lambda f, **kw: _Unpickler(f, **kw).load()


def save_model(path: str, inr: INR, mask: Volume, args: Namespace) -> None:
    torch.save({"model": inr.state_dict(), "mask": mask, "args": args}, path)


def _adapt_mlp(flat: torch.Tensor, net) -> torch.Tensor:
    """tcnn pads the first layer's input to a multiple of 16; the kernels here use 32 | 64 input columns."""
    want = net.params.numel()
    if flat.numel() == want:
        return flat
    o0, k0 = net.layer_shapes[0]
    rest = sum(o * k for o, k in net.layer_shapes[1:])
    k_ref = (flat.numel() - rest) // o0
    if k_ref <= 0 or k_ref > k0 or o0 * k_ref + rest != flat.numel():
        raise ValueError(f"checkpoint MLP has {flat.numel()} parameters, this build expects {want}")
    w0 = torch.zeros(o0, k0, dtype=flat.dtype, device=flat.device)
    w0[:, :k_ref] = flat[: o0 * k_ref].view(o0, k_ref)
    return torch.cat([w0.reshape(-1), flat[o0 * k_ref :]])


def load_model(path: str, device, args: Optional[Namespace] = None) -> Tuple[INR, Volume, Namespace]:
    """-> (INR with the checkpoint's parameters, mask Volume, the checkpoint's args overridden by `args`)."""
    cp = torch.load(path, map_location=device, weights_only=False, pickle_module=_pickle_module)
    cp_args = cp["args"]
    merged = Namespace(**vars(cp_args))
    if args is not None:
        for k, v in vars(args).items():
            setattr(merged, k, v)
    merged.device = device
    if not hasattr(merged, "dtype"):
        merged.dtype = torch.float32 if getattr(merged, "single_precision", False) else torch.float16
    state = dict(cp["model"])
    inr = INR(state["bounding_box"].to(device), merged)
    for name, net in (("density_net", inr.density_net),):
        key = f"{name}.params"
        if key in state and hasattr(net, "layer_shapes"):
            state[key] = _adapt_mlp(state[key].float(), net)
    state["encoding.params"] = state["encoding.params"].float()
    inr.load_state_dict(state)
    return inr.to(device), cp["mask"], merged
