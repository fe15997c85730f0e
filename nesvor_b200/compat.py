"""Zero-edit drop-in: run the UNMODIFIED reference package (daviddmc/NeSVoR) on libnesvor_b200.

The reference reaches native code (and, for files, one third-party package) through these imports:

* `import nesvor.slice_acq_cuda`          (nesvor/slice_acquisition/slice_acq.py:5-19; else it JIT-compiles its .cu files)
* `import nesvor.transform_convert_cuda`  (nesvor/transform/transform_convert.py:3-18; same fallback)
* `import tinycudann as tcnn`             (nesvor/nesvor/models.py:7; `tcnn.Encoding` :25, `tcnn.Network` :31)

* `import nibabel as nib`                 (nesvor/image/image.py:5; `nib.load`, `nib.nifti1.Nifti1Image`, `nib.save` only)

`install()` registers modules under those names whose functions / classes have the pybind / tcnn signatures and
call the C ABI (include/nesvor_b200.h) -- `forward / backward / adjoint_forward / adjoint_backward`,
`axisangle2mat_{forward,backward} / mat2axisangle_{forward,backward}` (each returning a list of tensors, absent masks as
empty tensors, like slice_acq_cuda.cpp:61-161 and transform_convert_cuda.cpp:27-69), `Encoding(n_input_dims,
encoding_config, dtype)` and `Network(n_input_dims, n_output_dims, network_config)`.  After

    import nesvor_b200.compat as compat; compat.install()
    import nesvor                                   # the reference, untouched

the reference's own `INR`, `NeSVoR`, `slice_acquisition`, `RigidTransform`, SRR ... run on the B200 kernels, one native op
per reference op.  For the hot loop that is not enough (one launch per op: 5-6 ms per iteration against 0.8 ms), so
`install(fused=True)` (the default) ALSO rebinds `nesvor.nesvor.train.train` -- what `nesvor reconstruct` calls
(nesvor/cli/commands.py:111) -- to `fused_train` below: the reference's own `Dataset` and `NeSVoR` classes, the reference's
loop structure (train.py:123-232), but each iteration is ONE launch of kernel A + the fused AdamW (`FusedTrainer.step`); the
INR it returns renders through `nsv_inr_render` underneath the reference's unmodified `sample_volume / sample_slices`.
Configurations kernel A is not instantiated for fall back to the reference's own `train` on the per-op path, with a warning.
A real `tinycudann` / `nibabel`, if installed, is left alone unless `tcnn="force"` / `nibabel="force"`; the nibabel stand-in
covers single-file NIfTI-1 (`image/nifti.py`).
"""
import importlib
import importlib.util
import sys
import types

import torch

_NAMES = ("nesvor.slice_acq_cuda", "nesvor.transform_convert_cuda", "tinycudann", "nibabel", "nibabel.nifti1")


def _lenient(fn):
    """The pybind modules insist on contiguous tensors (CHECK_CONTIGUOUS) and the reference's converter Functions hand
    autograd's gradients over as they come; newer torch versions produce expanded / strided gradients in places where the
    torch the reference was written for did not, so the stand-ins make tensor arguments contiguous instead of raising."""
    def call(*args):
        return fn(*[a.contiguous() if isinstance(a, torch.Tensor) and a.numel() and not a.is_contiguous() else a for a in args])

    call.__name__, call.__doc__, call.__module__ = fn.__name__, fn.__doc__, fn.__module__
    return call


def _slice_acq_module() -> types.ModuleType:
    sa = importlib.import_module("nesvor_b200.slice_acquisition.slice_acq")
    m = types.ModuleType("nesvor.slice_acq_cuda", "nesvor_b200 stand-in for the reference's slice_acq_cuda pybind module")
    for fn in ("forward", "backward", "adjoint_forward", "adjoint_backward"):
        setattr(m, fn, _lenient(getattr(sa, fn)))
    return m


def _transform_convert_module() -> types.ModuleType:
    tc = importlib.import_module("nesvor_b200.transform.transform_convert")
    m = types.ModuleType("nesvor.transform_convert_cuda", "nesvor_b200 stand-in for the reference's transform_convert_cuda pybind module")
    for fn in ("axisangle2mat_forward", "axisangle2mat_backward", "mat2axisangle_forward", "mat2axisangle_backward"):
        setattr(m, fn, _lenient(getattr(tc, fn)))
    return m


def _tcnn_module() -> types.ModuleType:
    from .nesvor.encoding import FusedMLP, HashGridEncoding

    class Encoding(HashGridEncoding):
        """tcnn.Encoding(n_input_dims, encoding_config, dtype=torch.float16): the HashGrid otype NeSVoR uses."""

        def __init__(self, n_input_dims, encoding_config, dtype=torch.float16, seed=1337):
            super().__init__(n_input_dims, dict(encoding_config), dtype, seed)

    class Network(FusedMLP):
        """tcnn.Network(n_input_dims, n_output_dims, network_config): ReLU MLP without biases, fp16 tensor cores."""

        def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
            super().__init__(n_input_dims, n_output_dims, dict(network_config), seed)

    m = types.ModuleType("tinycudann", "nesvor_b200 stand-in for the two tiny-cuda-nn modules NeSVoR instantiates")
    m.Encoding, m.Network = Encoding, Network
    m.__nesvor_b200_shim__ = True
    return m


def _nibabel_module() -> types.ModuleType:
    """The three nibabel calls the reference makes (nesvor/image/image.py:253-296) on image/nifti.py:
    `nib.load(path)` -> object with `.header["dim" | "pixdim"]`, `.get_fdata()`, `.affine`, `.get_qform()`;
    `nib.nifti1.Nifti1Image(data, affine)` with `.header.set_xyzt_units / set_qform / set_sform`; `nib.save(img, path)`."""
    import numpy as np

    from .image.nifti import read_nifti, write_nifti

    class _Header(dict):
        def set_xyzt_units(self, *a, **k):  # written files always carry millimetres
            pass

        def set_qform(self, affine, code=None):  # written files carry the affine as qform "aligned" ...
            self["_affine"] = np.asarray(affine, np.float64)

        def set_sform(self, affine, code=None):  # ... and as sform "scanner"
            self["_affine"] = np.asarray(affine, np.float64)

    class Nifti1Image:
        def __init__(self, dataobj, affine, header=None):
            self._data = np.asarray(dataobj)
            self.affine = np.eye(4) if affine is None else np.asarray(affine, np.float64)
            self.header = _Header(header or {})
            self._qform = None

        def get_fdata(self):
            return np.asarray(self._data, np.float64)

        def get_qform(self):
            return self.affine if self._qform is None else self._qform

    def load(path):
        data, hdr = read_nifti(str(path))
        img = Nifti1Image(data, hdr["affine"], {"dim": hdr["dim"], "pixdim": hdr["pixdim"].astype(np.float32)})
        img._qform = hdr["qform"]
        return img

    def save(img, path):
        write_nifti(str(path), img._data, img.header.get("_affine", img.affine))

    m = types.ModuleType("nibabel", "nesvor_b200 stand-in for the three nibabel calls NeSVoR makes (NIfTI-1 single files)")
    n1 = types.ModuleType("nibabel.nifti1")
    n1.Nifti1Image = Nifti1Image
    m.nifti1, m.Nifti1Image, m.load, m.save = n1, Nifti1Image, load, save
    m.__nesvor_b200_shim__ = n1.__nesvor_b200_shim__ = True
    return m


def _have_real(name: str) -> bool:
    cur = sys.modules.get(name)
    if cur is not None:
        return not getattr(cur, "__nesvor_b200_shim__", False)
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


_REF_TRAIN = None  # the reference's own train(), kept for configurations outside kernel A's instantiations
LAST_TRAIN_INFO: dict = {}  # what the last fused_train call did (path taken, ms per iteration): read by tools / tests


class _DeferredBatch:
    """What the fused INR's `sample_batch` hands to its `forward`: the un-expanded request (sample.py:25-31,44-50)."""

    def __init__(self, xyz, transformation, psf_sigma, n_samples):
        self.xyz, self.transformation, self.psf_sigma, self.n_samples = xyz, transformation, psf_sigma, n_samples


def _attach_fused_render(inr, args) -> None:
    """Makes `inr.sample_batch(...)` + `inr(batch, False).mean(-1)` -- the two calls of the reference's sample.py:25-31 and
    :44-50 -- one launch of the forward-only fused kernel `nsv_inr_render`: sample_batch defers, forward renders and returns
    [M, 1] so that the caller's `.mean(-1)` is the identity.  Plain tensors still take the module's own forward."""
    from .nesvor.fused import attach_render_state, fused_render

    attach_render_state(inr, args)
    plain_forward = inr.forward

    def sample_batch(xyz, transformation, psf_sigma, n_samples):
        return _DeferredBatch(xyz, transformation, psf_sigma, n_samples)

    def forward(x, return_all=True):
        if isinstance(x, _DeferredBatch):
            return fused_render(inr, x.xyz, x.transformation, x.psf_sigma, x.n_samples)[:, None]
        return plain_forward(x, return_all)

    inr.sample_batch, inr.forward = sample_batch, forward


def fused_train(slices, args):
    """Drop-in for the reference's `train(slices, args) -> (INR, List[Slice], Volume)` (nesvor/nesvor/train.py:123-232).

    Same data path and bookkeeping, built from the REFERENCE'S OWN classes (its `Dataset`: pixel table, epoch shuffle,
    bounding box, mean, output mask; its `NeSVoR` on the tinycudann stand-ins; its `TrainLogger`), same optimiser
    hyper-parameters, milestones and LR decay -- but the body of the loop (train.py:183-197: autocast forward, scaled
    backward, AdamW step, zero_grad) is `FusedTrainer.step`: kernel A + finalize (+ transReg) + the fused AdamW.  The
    reference reads every loss back every iteration (`.item()`, train.py:199-200: 5-6 host syncs); here the bias-corrected
    moving average (utils/misc.py:91-122, alpha = 0.999) is kept on the device and read only when a log line is printed."""
    import datetime
    import logging
    import time

    import nesvor.nesvor.train as rt

    from .nesvor.fused import FusedTrainer, FusedUnsupported

    dataset = rt.Dataset(slices, args)
    model = rt.NeSVoR(dataset.transformation, dataset.resolution, dataset.mean, dataset.bounding_box, args)
    LAST_TRAIN_INFO.clear()
    try:
        if args.single_precision:
            raise FusedUnsupported("--single-precision selects the fp32 nn.Linear MLPs (with biases)")
        trainer = FusedTrainer(model, args)
        batch = dataset.get_batch(args.batch_size, args.device)
        first = trainer.step(**batch)  # the launch itself reports a configuration without an instantiation
    except FusedUnsupported as e:
        logging.warning("nesvor_b200: %s -- training on the per-op native path (reference loop)", e)
        LAST_TRAIN_INFO.update(path="reference loop on per-op native kernels", why=str(e))
        return _REF_TRAIN(slices, args)
    logging.debug(rt.log_params(model))
    decay_milestones = [int(m * args.n_iter) for m in args.milestones]
    model.train()
    keys = list(first.keys())
    alpha = 1 - 0.001
    # moving average (logger.py's) of the whole 8-float loss slot, one lerp kernel per iteration; the columns the log prints are
    # picked when a line is due (`first`'s values are views into the slot: their storage offsets name the columns)
    slot0 = trainer.state.losses
    cols = [first[k].storage_offset() - slot0.storage_offset() for k in keys]
    ema = torch.zeros(8, dtype=torch.float32, device=args.device)
    train_logger = None
    logging.info("NeSVoR training starts (nesvor_b200 fused iteration).")
    torch.cuda.synchronize(args.device)
    t_start = time.time()
    losses = first
    for i in range(1, args.n_iter + 1):
        if i > 1:
            batch = dataset.get_batch(args.batch_size, args.device)
            losses = trainer.step(**batch)
        ema.lerp_(trainer.state.losses, 1 - alpha)  # ema = alpha * ema + (1 - alpha) * losses
        if (decay_milestones and i >= decay_milestones[0]) or i == args.n_iter:
            if train_logger is None:
                train_logger = rt.TrainLogger("time", "epoch", "iter", *keys, "lr")
            full = (ema / (1 - alpha**i)).tolist()  # the only device read-back of the loop
            avg = [full[c] for c in cols]
            train_logger.log(datetime.timedelta(seconds=int(time.time() - t_start)), dataset.epoch, i, *avg, trainer.lr)
            if i < args.n_iter:
                decay_milestones.pop(0)
                trainer.decay_lr(args.gamma)
    torch.cuda.synchronize(args.device)
    wall = time.time() - t_start
    LAST_TRAIN_INFO.update(path="fused (kernel A + fused AdamW)", iterations=args.n_iter, ms_per_iteration=1e3 * wall / max(args.n_iter, 1),
                           queries_per_iteration=args.batch_size * args.n_samples, launches_per_iteration=4 if trainer.pose else 3,
                           n_levels=int(model.inr.encoding.n_levels), n_pixels=int(dataset.xyz.shape[0]))
    trainer.sync_to_model()
    _attach_fused_render(model.inr, args)
    # outputs (train.py:223-232)
    transformation = model.transformation
    dataset.transformation = transformation
    mask = dataset.mask
    output_slices = []
    for i in range(len(slices)):
        output_slice = slices[i].clone()
        output_slice.transformation = transformation[i]
        output_slices.append(output_slice)
    return model.inr, output_slices, mask


def _install_fused_train() -> None:
    """Rebinds `train` where the reference looks it up: the defining module and, if already imported, the command module
    that did `from ..nesvor.train import train` (nesvor/cli/commands.py)."""
    global _REF_TRAIN
    rt = importlib.import_module("nesvor.nesvor.train")
    if getattr(rt.train, "__nesvor_b200_fused__", False):
        return
    _REF_TRAIN = rt.train
    fused_train.__nesvor_b200_fused__ = True
    rt.train = fused_train
    pkg = sys.modules.get("nesvor.nesvor")
    if pkg is not None and getattr(pkg, "train", None) is _REF_TRAIN:
        pkg.train = fused_train
    cmd = sys.modules.get("nesvor.cli.commands")
    if cmd is not None and getattr(cmd, "train", None) is _REF_TRAIN:
        cmd.train = fused_train


def install(tcnn: str = "auto", nibabel: str = "auto", fused: bool = True) -> dict:
    """Registers the stand-in modules in `sys.modules`; returns {name: module} of what was installed.
    tcnn / nibabel = "auto": only when no real `tinycudann` / `nibabel` is importable; "force": always; "never": leave it alone.
    fused: also rebind the reference's `train` to `fused_train` (needs the reference package `nesvor` on sys.path; silently
    skipped when it is not importable yet -- call `install()` again after adding it)."""
    done = {}
    for name, make in (("nesvor.slice_acq_cuda", _slice_acq_module), ("nesvor.transform_convert_cuda", _transform_convert_module)):
        sys.modules[name] = done[name] = make()
    if tcnn == "force" or (tcnn == "auto" and not _have_real("tinycudann")):
        sys.modules["tinycudann"] = done["tinycudann"] = _tcnn_module()
    if nibabel == "force" or (nibabel == "auto" and not _have_real("nibabel")):
        nib = _nibabel_module()
        sys.modules["nibabel"], sys.modules["nibabel.nifti1"] = nib, nib.nifti1
        done["nibabel"] = nib
    parent = sys.modules.get("nesvor")  # `import nesvor.x as y` resolves through the parent's attribute first
    if parent is not None:
        for name in ("slice_acq_cuda", "transform_convert_cuda"):
            setattr(parent, name, sys.modules["nesvor." + name])
    if fused:
        try:
            _install_fused_train()
            done["nesvor.nesvor.train.train"] = fused_train
        except ImportError:
            pass
    return done


def uninstall() -> None:
    for name in _NAMES:
        mod = sys.modules.get(name)
        if mod is not None and (not name.startswith(("tinycudann", "nibabel")) or getattr(mod, "__nesvor_b200_shim__", False)):
            del sys.modules[name]
