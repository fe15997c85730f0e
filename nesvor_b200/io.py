"""Model checkpoints of the INR path: `model.pt` = {"model": INR.state_dict(), "mask": Volume, "args": Namespace}
(nesvor/cli/io.py:36-59).

`save_model` writes exactly that dictionary.  `load_model` also reads checkpoints written by the reference: its
pickled helper objects (`nesvor.image.image.Volume`, `nesvor.transform.transform.RigidTransform`) are mapped onto this
package's classes (same attribute names), and tiny-cuda-nn's flat parameters are taken over as they are -- the hash
table is the same level-major `[sum_l T_l, F]` fp32 vector and the MLP the same row-major `[out, in]` layers; only a
first layer whose input tcnn pads to 16 columns is widened to the 32 columns the kernels are instantiated for.

`inputs(args)` / `outputs(data, args)` mirror nesvor/cli/io.py:9-49: the argument-driven loading of stacks (+ masks,
thickness overrides), slice folders and a model checkpoint, and the saving of the output volume (optionally rescaled),
model, motion-corrected slices and simulated slices -- NIfTI through nesvor_b200.image (no nibabel).
"""
import pickle
import types
from argparse import Namespace
from typing import Any, Dict, Optional, Tuple

import torch

from .image import Volume, load_slices, load_stack, save_slices
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


def _export_mlp(flat: torch.Tensor, net) -> torch.Tensor:
    """Inverse of `_adapt_mlp`: the kernels keep the first layer with 32 | 64 input columns, tcnn with the
    input width padded to a multiple of 16; the extra columns only ever multiply zeros and are dropped on export, so that
    `density_net.params` has tcnn's element count and layout."""
    if not hasattr(net, "layer_shapes"):
        return flat
    o0, k0 = net.layer_shapes[0]
    k_ref = (net.n_input_dims + 15) // 16 * 16
    if k_ref >= k0:
        return flat
    w0 = flat[: o0 * k0].view(o0, k0)[:, :k_ref]
    return torch.cat([w0.reshape(-1), flat[o0 * k0 :]])


def save_model(path: str, inr: INR, mask: Volume, args: Namespace) -> None:
    """cli/io.py:36-46: {"model": state_dict, "mask": Volume, "args": Namespace}; the MLP parameters are written in tcnn's
    16-padded layout (`_export_mlp`), which `load_model` widens again."""
    state = dict(inr.state_dict())
    if "density_net.params" in state:
        state["density_net.params"] = _export_mlp(state["density_net.params"], inr.density_net)
    torch.save({"model": state, "mask": mask, "args": args}, path)


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
    """-> (INR with the checkpoint's parameters, mask Volume, the checkpoint's args overridden by `args`).

    Like the reference (cli/io.py:24-29,53-59) the network is built STRICTLY from the checkpoint's own args -- the caller's
    namespace (which `inputs()` passes in full, model hyper-parameters included) must never change the architecture the
    stored parameters are decoded with -- and the two namespaces are merged only afterwards, for downstream use."""
    cp = torch.load(path, map_location=device, weights_only=False, pickle_module=_pickle_module)
    cp_args = cp["args"]
    build = Namespace(**vars(cp_args))
    build.device = device
    if not hasattr(build, "dtype"):
        build.dtype = torch.float32 if getattr(build, "single_precision", False) else torch.float16
    state = dict(cp["model"])
    inr = INR(state["bounding_box"].to(device), build)
    for name, net in (("density_net", inr.density_net),):
        key = f"{name}.params"
        if key in state and hasattr(net, "layer_shapes"):
            state[key] = _adapt_mlp(state[key].float(), net)
    state["encoding.params"] = state["encoding.params"].float()
    inr.load_state_dict(state)
    merged = Namespace(**vars(cp_args))  # utils/misc.py:22-26 merge_args: the caller's values win
    if args is not None:
        for k, v in vars(args).items():
            setattr(merged, k, v)
    merged.device = device
    if not hasattr(merged, "dtype"):
        merged.dtype = build.dtype
    return inr.to(device), cp["mask"], merged


def inputs(args: Namespace) -> Tuple[Dict[str, Any], Namespace]:
    """cli/io.py:9-32: {"input_stacks": [Stack], "input_slices": [Slice], "model": INR, "mask": Volume} for whatever of
    `args.input_stacks` (+ `stack_masks`, `thicknesses`), `args.input_slices`, `args.input_model` is set; with a model the
    checkpoint's args are merged under the caller's."""
    out: Dict[str, Any] = {}
    if getattr(args, "input_stacks", None) is not None:
        masks, thick = getattr(args, "stack_masks", None), getattr(args, "thicknesses", None)
        out["input_stacks"] = []
        for i, f in enumerate(args.input_stacks):
            stack = load_stack(f, masks[i] if masks is not None else None, device=args.device)
            if thick is not None:
                stack.thickness = thick[i]
            out["input_stacks"].append(stack)
    if getattr(args, "input_slices", None) is not None:
        out["input_slices"] = load_slices(args.input_slices, args.device)
    if getattr(args, "input_model", None) is not None:
        out["model"], out["mask"], args = load_model(args.input_model, args.device, args)
    return out, args


def outputs(data: Dict[str, Any], args: Namespace) -> None:
    """cli/io.py:35-49: writes `output_volume` (rescaled to `output_intensity_mean` when set), `output_model`,
    `output_slices` and `simulated_slices` for the keys present in `data` and named in `args`."""
    if getattr(args, "output_volume", None) and "output_volume" in data:
        if getattr(args, "output_intensity_mean", None):
            data["output_volume"].rescale(args.output_intensity_mean)
        data["output_volume"].save(args.output_volume)
    if getattr(args, "output_model", None) and "output_model" in data:
        save_model(args.output_model, data["output_model"], data["mask"], args)
    if getattr(args, "output_slices", None) and "output_slices" in data:
        save_slices(args.output_slices, data["output_slices"])
    if getattr(args, "simulated_slices", None) and "simulated_slices" in data:
        save_slices(args.simulated_slices, data["simulated_slices"])
