"""Host side of the fused path: packs a `NeSVoR` model into the flat buffers kernel A reads, runs
`nsv_inr_train_step` + `nsv_adamw_step` per iteration, and renders with `nsv_inr_render`.

What it replaces in the reference: the body of the hot loop (nesvor/nesvor/train.py:179-198:
autocast forward, GradScaler-scaled backward, AdamW step, zero_grad) and the inference batches of
nesvor/nesvor/sample.py:23-32,44-50.  The learnable tensors live in ONE flat fp32 buffer

    [ hash table | packed MLP weights | slice_embedding | logit_coef | log_var_slice | axisangle ]

(trainable prefix first) with a parallel gradient buffer, Adam moments and an fp16 shadow of the
table + weights, so that an iteration is: 1 fused forward/backward launch (+ a 1-block finalize),
1 AdamW launch over the whole trainable prefix, and -- only when poses are optimised -- the
batch-independent `transReg` term (models.py:357-363) through the native pose converters.
"""
import ctypes
import math
from argparse import Namespace
from typing import Dict, Optional

import torch

from .. import _lib
from ..transform import RigidTransform
from .encoding import FusedMLP, HashGridEncoding
from .models import B_REG, D_LOSS, DS_LOSS, I_REG, INR, S_LOSS, T_REG, NeSVoR

_IMAGE_REG = {"TV": 1, "edge": 2, "L2": 3}
_SIGMA_Z_SLOT = 16  # packed sigma_net input = [slice embedding (16) | z0..z15]; z0's column is dead


LOSS_RING = 4096  # iterations before a loss slot is reused (FusedState.next_losses)


class FusedUnsupported(RuntimeError):
    """The configuration is outside what kernel A is instantiated for (use the unfused native path)."""


def _require(cond: bool, why: str) -> None:
    if not cond:
        raise FusedUnsupported("fused INR path unsupported: " + why)


def make_config(inr: INR, args: Namespace, *, delta: float = 0.0, n_batch_samples: int = 1 << 20,
                pixel_variance=False, slice_variance=False, slice_scale=False, pose_grad=False) -> _lib.InrConfig:
    enc = inr.encoding
    _require(isinstance(enc, HashGridEncoding) and isinstance(inr.density_net, FusedMLP),
             "model must be built with dtype=float16 (tcnn-style modules); the fp32 nn.Linear branch has biases")
    _require(enc.n_features_per_level == 2 and enc.n_levels <= 16, "needs n_features_per_level=2 and n_levels<=16")
    _require(args.width in (32, 64) and 1 <= args.depth <= 3, "needs width in {32,64} and depth in 1..3")
    _require(inr.density_net.n_out_padded == 16, "needs 1 + n_features_z <= 16")
    cfg = _lib.InrConfig()
    cfg.grid = enc.meta
    cfg.width, cfg.depth = args.width, args.depth
    cfg.n_features_z, cfg.n_features_slice = args.n_features_z, args.n_features_slice
    cfg.n_levels_bias = args.n_levels_bias
    cfg.pixel_variance, cfg.slice_variance = int(pixel_variance), int(slice_variance)
    cfg.slice_scale, cfg.pose_grad = int(slice_scale), int(pose_grad)
    cfg.image_reg = _IMAGE_REG[args.image_regularization] if getattr(args, "weight_image", 0) else 0
    cfg.delta = float(delta)
    cfg.w_image = float(getattr(args, "weight_image", 0.0))
    cfg.w_bias = float(getattr(args, "weight_bias", 0.0))
    bb = inr.bounding_box.detach().float().cpu()
    for i in range(3):
        cfg.bbox_lo[i] = float(bb[0, i])
        cfg.bbox_hi[i] = float(bb[1, i])
    # loss scale for the fp16 backward operands: dL/dz ~ 1/(B*S); power of two, so unscaling is exact
    cfg.grad_scale = float(2.0 ** (math.floor(math.log2(max(n_batch_samples, 1))) + 4))
    return cfg


def _mlp_layout(cfg: _lib.InrConfig):
    off = (ctypes.c_int64 * 3)()
    total = _lib.lib().nsv_inr_mlp_layout(ctypes.byref(cfg), off)
    if total < 0:
        _lib.check(int(total), "nsv_inr_mlp_layout")
    return int(total), [int(o) for o in off]


def _pack_sigma(logical: torch.Tensor, width: int) -> torch.Tensor:
    """FusedMLP flat params of sigma_net -> packed layout (first layer columns re-slotted)."""
    w0 = logical[: width * 32].view(width, 32)
    p0 = torch.zeros_like(w0)
    p0[:, :16] = w0[:, :16]
    p0[:, _SIGMA_Z_SLOT + 1 : 32] = w0[:, 16:31]
    return torch.cat([p0.reshape(-1), logical[width * 32 :]])


def _unpack_sigma(packed: torch.Tensor, width: int) -> torch.Tensor:
    p0 = packed[: width * 32].view(width, 32)
    w0 = torch.zeros_like(p0)
    w0[:, :16] = p0[:, :16]
    w0[:, 16:31] = p0[:, _SIGMA_Z_SLOT + 1 : 32]
    return torch.cat([w0.reshape(-1), packed[width * 32 :]])


class FusedState:
    """Flat parameter / gradient buffers of one NeSVoR model (or of a bare INR for rendering)."""

    def __init__(self, inr: INR, args: Namespace, model: Optional[NeSVoR] = None, n_batch_samples: int = 1 << 20):
        self.inr, self.model, self.args = inr, model, args
        dev = inr.encoding.params.device
        self.device = dev
        a = args
        pv = model is not None and not a.no_pixel_variance
        sv = model is not None and not a.no_slice_variance
        sc = model is not None and not a.no_slice_scale
        pg = model is not None and not a.no_transformation_optimization
        if model is not None:
            _require(a.n_levels_bias == 0 or (pv and 1 <= a.n_levels_bias <= 4 and a.width == 64),
                     "the fused bias-field head needs 1 <= n_levels_bias <= 4, width 64 and the pixel-variance (sigma_net) heads on")
            _require(not pv or (a.depth == 1 and a.n_features_slice == 16 and a.n_features_z == 15),
                     "fused sigma_net needs depth=1, n_features_slice=16, n_features_z=15")
        self.cfg = make_config(inr, a, delta=(model.delta if model is not None else 0.0), n_batch_samples=n_batch_samples,
                               pixel_variance=pv, slice_variance=sv, slice_scale=sc, pose_grad=pg)
        n_mlp, (self.off_density, self.off_sigma, self.off_bias) = _mlp_layout(self.cfg)
        n_table = inr.encoding.params.numel()
        ns = model.n_slices if model is not None else 0
        # ---- segments: (name, numel, trainable) in buffer order, trainable prefix first ----
        segs = [("table", n_table, True), ("mlp", n_mlp, True)]
        if model is not None:
            tail = []
            for name, n, used in (("slice_embedding", ns * a.n_features_slice if a.n_features_slice else 0, pv),
                                  ("logit_coef", ns if sc else 0, sc), ("log_var_slice", ns if sv else 0, sv),
                                  ("axisangle", ns * 6, pg)):
                if n == 0:
                    continue
                (segs if used else tail).append((name, n, used))
            segs += tail
        self.offsets: Dict[str, slice] = {}
        off = 0
        self.n_train = 0
        for name, n, train in segs:
            n_pad = (n + 3) // 4 * 4  # keep every segment 16-byte aligned
            self.offsets[name] = slice(off, off + n)
            off += n_pad
            if train:
                self.n_train = off
        self.n_total = off
        # per-slice parameters (everything behind the MLP weights) are read by kernel A in fp32: under the peer-memory
        # optimiser every rank keeps a current fp32 mirror of them (`tail32`, see enable_peer_memory / seg)
        self.tail_lo = (self.offsets["mlp"].stop + 3) // 4 * 4
        self.tail32: Optional[torch.Tensor] = None
        self.mc_ptrs = None  # multicast addresses of (gradient, fp16 copy, fp32 tail mirror) once peer memory is on
        self.flat = torch.zeros(self.n_total, dtype=torch.float32, device=dev)
        self.flat16 = torch.zeros(self.n_total, dtype=torch.float16, device=dev)
        self.grad = torch.zeros(self.n_total + 8, dtype=torch.float32, device=dev)
        # loss values: a ring of 8-float slots, one per iteration (`next_losses`), so that an iteration needs neither a memset of
        # its slot nor a snapshot copy of it -- the whole ring is cleared once per LOSS_RING iterations
        self.loss_ring = torch.zeros(LOSS_RING, 8, dtype=torch.float32, device=dev)
        self.loss_slot = 0
        self.losses = self.loss_ring[0]
        self.last_readback: Optional[torch.cuda.Event] = None
        self.psf_sigma = model.psf_sigma.contiguous().float() if model is not None else None
        self.pull_from_model()

    # ------------------------------------------------------------------ data-parallel peer memory
    def enable_peer_memory(self, group) -> None:
        """Moves the gradient buffer and the fp16 parameter copy into ONE symmetric allocation (CUDA peer memory set up
        by torch.distributed._symmetric_memory over NVLink) and records every rank's pointers, so that
        nsv_adamw_step_dp can read all ranks' gradients and write all ranks' fp16 parameters directly."""
        import torch.distributed._symmetric_memory as symm_mem

        n_grad = self.grad.numel() * 4
        n_grad_pad = (n_grad + 255) // 256 * 256
        n_f16_pad = (self.flat16.numel() * 2 + 255) // 256 * 256
        n_tail = self.n_total - self.tail_lo
        n_tail_pad = (n_tail * 4 + 255) // 256 * 256
        nbytes = n_grad_pad + n_f16_pad + n_tail_pad + 512  # + 64 x u64 rendezvous flags (nsv_adamw_step_dp_sync)
        buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
        handle = symm_mem.rendezvous(buf, group)
        grad = buf[:n_grad].view(torch.float32)
        flat16 = buf[n_grad_pad : n_grad_pad + self.flat16.numel() * 2].view(torch.float16)
        grad.zero_()
        flat16.copy_(self.flat16)
        self.grad, self.flat16 = grad, flat16
        if n_tail > 0:  # only the owner of a shard updates its fp32 master: the owner mirrors the per-slice parameters everywhere
            tail32 = buf[n_grad_pad + n_f16_pad : n_grad_pad + n_f16_pad + n_tail * 4].view(torch.float32)
            tail32.copy_(self.flat[self.tail_lo :])
            self.tail32 = tail32
        self._symm_buf, self.peer_handle = buf, handle
        world = handle.world_size
        ptrs = [int(p) for p in handle.buffer_ptrs]
        self.peer_grads = (ctypes.c_void_p * world)(*ptrs)
        self.peer_flat16 = (ctypes.c_void_p * world)(*[p + n_grad_pad for p in ptrs])
        self.peer_tail32 = (ctypes.c_void_p * world)(*[p + n_grad_pad + n_f16_pad for p in ptrs]) if n_tail > 0 else None
        self.dp_flags = buf[n_grad_pad + n_f16_pad + n_tail_pad :].view(torch.int64)
        self.dp_flags.zero_()
        self.peer_flags = (ctypes.c_void_p * world)(*[p + n_grad_pad + n_f16_pad + n_tail_pad for p in ptrs])
        # NVSwitch multicast address of the same allocation (0 when the fabric / driver has no multicast support): lets the
        # optimiser kernel sum the gradients and replicate the parameters inside the switch (nsv_adamw_step_dp_mc)
        mc = int(getattr(handle, "multicast_ptr", 0) or 0)
        self.mc_ptrs = (mc, mc + n_grad_pad, (mc + n_grad_pad + n_f16_pad) if n_tail > 0 else 0) if mc else None
        torch.cuda.synchronize(self.device)
        handle.barrier()  # every rank's copy is initialised before anyone's optimiser writes into it

    def disable_peer_memory(self) -> None:
        """Back to private buffers (another rank could not set peer memory up: the ranks fall back together)."""
        self.grad, self.flat16 = self.grad.clone(), self.flat16.clone()
        self.tail32 = None
        self._symm_buf = self.peer_handle = self.peer_grads = self.peer_flat16 = self.peer_tail32 = self.peer_flags = self.dp_flags = None
        self.mc_ptrs = None

    # ------------------------------------------------------------------ model <-> flat
    def seg(self, name: str, buf: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        if name not in self.offsets:
            return None
        sl = self.offsets[name]
        if buf is None and self.tail32 is not None and sl.start >= self.tail_lo:
            return self.tail32[sl.start - self.tail_lo : sl.stop - self.tail_lo]
        return (self.flat if buf is None else buf)[sl]

    def pull_from_model(self) -> None:
        with torch.no_grad():
            self.seg("table").copy_(self.inr.encoding.params)
            mlp = self.seg("mlp")
            d = self.inr.density_net.params
            mlp[self.off_density : self.off_density + d.numel()].copy_(d)
            m = self.model
            if m is not None:
                if self.cfg.pixel_variance:
                    s = _pack_sigma(m.sigma_net.params.detach(), self.args.width)
                    mlp[self.off_sigma : self.off_sigma + s.numel()].copy_(s)
                if self.cfg.n_levels_bias:  # b_net's logical column order [slice embedding | pe_bias | pad] is the packed one
                    b = m.b_net.params.detach()
                    mlp[self.off_bias : self.off_bias + b.numel()].copy_(b)
                for name, t in (("slice_embedding", getattr(getattr(m, "slice_embedding", None), "weight", None)),
                                ("logit_coef", getattr(m, "logit_coef", None)), ("log_var_slice", getattr(m, "log_var_slice", None)),
                                ("axisangle", m.axisangle)):
                    if name in self.offsets and t is not None:
                        self.seg(name).copy_(t.reshape(-1))
            self.flat16.copy_(self.flat)

    def push_to_model(self) -> None:
        with torch.no_grad():
            self.inr.encoding.params.copy_(self.seg("table"))
            mlp = self.seg("mlp")
            d = self.inr.density_net.params
            d.copy_(mlp[self.off_density : self.off_density + d.numel()])
            m = self.model
            if m is not None:
                if self.cfg.pixel_variance:
                    n = m.sigma_net.params.numel()
                    m.sigma_net.params.copy_(_unpack_sigma(mlp[self.off_sigma : self.off_sigma + n], self.args.width))
                if self.cfg.n_levels_bias:
                    n = m.b_net.params.numel()
                    m.b_net.params.copy_(mlp[self.off_bias : self.off_bias + n])
                for name, t in (("slice_embedding", getattr(getattr(m, "slice_embedding", None), "weight", None)),
                                ("logit_coef", getattr(m, "logit_coef", None)), ("log_var_slice", getattr(m, "log_var_slice", None)),
                                ("axisangle", m.axisangle)):
                    if name in self.offsets and t is not None:
                        t.copy_(self.seg(name).view_as(t))

    # ------------------------------------------------------------------ native structs
    def params_struct(self) -> _lib.InrParams:
        p = _lib.InrParams()
        p.table_f16 = self.seg("table", self.flat16).data_ptr()
        p.mlp_f16 = self.seg("mlp", self.flat16).data_ptr()
        for name in ("axisangle", "slice_embedding", "logit_coef", "log_var_slice"):
            t = self.seg(name)
            setattr(p, name, t.data_ptr() if t is not None and t.numel() else None)
        p.psf_sigma = self.psf_sigma.data_ptr() if self.psf_sigma is not None else None
        p.n_slices = self.model.n_slices if self.model is not None else 0
        return p

    def grads_struct(self) -> _lib.InrGrads:
        g = _lib.InrGrads()
        g.table = self.seg("table", self.grad).data_ptr()
        g.mlp = self.seg("mlp", self.grad).data_ptr()
        for field, name in (("axisangle", "axisangle"), ("slice_embedding", "slice_embedding"), ("slice_scale_c", "logit_coef"),
                            ("log_var_slice", "log_var_slice")):
            t = self.seg(name, self.grad)
            setattr(g, field, t.data_ptr() if t is not None and t.numel() else None)
        g.losses = self.losses.data_ptr()
        return g

    # ------------------------------------------------------------------ one forward+backward
    def forward_backward(self, xyz, v, slice_idx, noise=None, seed: int = 0, offset: int = 0, want_v_out: bool = False,
                         dist=None, world: int = 1):
        """Accumulates gradients into `self.grad` (caller zeroes, losses included) and returns (losses[8] view, v_out or None).
        With the bias-field head, `nsv_inr_bias_mean` first leaves mean(log_bias) of the batch in losses[4] (biasReg couples
        all samples, models.py:323); data-parallel callers pass `dist` / `world` so that the mean is the global one."""
        B = xyz.shape[0]
        S = self.args.n_samples
        xyz = xyz.contiguous().float()
        v = v.contiguous().float()
        slice_idx = slice_idx.contiguous().to(torch.int64)
        if noise is not None:
            noise = noise.contiguous().float()
            assert noise.shape == (B, S, 3)
        v_out = torch.empty(B, dtype=torch.float32, device=xyz.device) if want_v_out else None
        prm, grd = self.params_struct(), self.grads_struct()
        if self.cfg.n_levels_bias:
            with torch.cuda.device(self.device):
                rc = _lib.lib().nsv_inr_bias_mean(
                    ctypes.byref(self.cfg), ctypes.byref(prm), _lib.ptr(xyz), _lib.ptr(slice_idx), _lib.ptr(noise), ctypes.c_uint64(seed),
                    ctypes.c_uint64(offset), ctypes.c_void_p(self.losses.data_ptr() + 16), ctypes.c_int64(B), ctypes.c_int(S),
                    _lib.stream(self.device))
            if rc == -2:
                raise FusedUnsupported(_lib.lib().nsv_last_error_string().decode())
            _lib.check(rc, "nsv_inr_bias_mean")
            if dist is not None and world > 1:  # equal per-rank batch sizes: the global mean is the mean of the ranks' means
                m = self.losses[4:5]
                dist.all_reduce(m)
                m.div_(world)
        with torch.cuda.device(self.device):
            rc = _lib.lib().nsv_inr_train_step(
                ctypes.byref(self.cfg), ctypes.byref(prm), ctypes.byref(grd), _lib.ptr(xyz), _lib.ptr(v), _lib.ptr(slice_idx),
                _lib.ptr(noise), ctypes.c_uint64(seed), ctypes.c_uint64(offset), _lib.ptr(v_out), ctypes.c_int64(B), ctypes.c_int(S),
                _lib.stream(self.device))
        if rc == -2:
            raise FusedUnsupported(_lib.lib().nsv_last_error_string().decode())
        _lib.check(rc, "nsv_inr_train_step")
        if self.cfg.pixel_variance:  # the z0 slot of sigma_net's first layer is structurally zero
            w = self.args.width
            g0 = self.seg("mlp", self.grad)[self.off_sigma : self.off_sigma + w * 32].view(w, 32)
            g0[:, _SIGMA_Z_SLOT] = 0
        return self.losses, v_out

    def next_losses(self) -> torch.Tensor:
        """Binds `self.losses` to the next (zeroed) slot of the ring and returns it.  Slots handed out earlier keep their values
        until the ring wraps, LOSS_RING - 1 iterations later -- long enough for every consumer here (train(), compat.fused_train
        and the benchmark read an iteration's values one iteration later)."""
        self.loss_slot += 1
        if self.loss_slot == self.loss_ring.shape[0]:
            if self.last_readback is not None:  # a read-back stream may still be copying the newest slot (losses_to_host)
                torch.cuda.current_stream(self.device).wait_event(self.last_readback)
            self.loss_ring.zero_()
            self.loss_slot = 0
        self.losses = self.loss_ring[self.loss_slot]
        return self.losses

    def loss_dict(self, losses: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Views into `losses` (no device work) under the reference's names; [6] = MSE + logVar is written by the finalize kernel."""
        a = self.args
        out = {D_LOSS: losses[0]}
        if not (a.no_pixel_variance and a.no_slice_variance):
            out[S_LOSS] = losses[1]
            out[DS_LOSS] = losses[6]
        if a.n_levels_bias:
            out[B_REG] = losses[2]
        out[I_REG] = losses[3]
        return out


class LossHandle:
    """The loss values of one iteration on their way to the host: a pinned buffer filled by an asynchronous copy that was
    enqueued right behind the iteration's kernels, plus the event that marks its completion.  `get()` blocks only until
    THAT copy has landed, so a loop that reads iteration i's losses after enqueueing iteration i + 1 (train() below does)
    keeps one iteration of work queued on the GPU instead of draining it at every `.item()` (train.py:199-200)."""

    def __init__(self, keys, host: torch.Tensor, event: torch.cuda.Event, pos=None):
        self.keys, self.host, self.event = keys, host, event
        self.pos = list(range(len(keys))) if pos is None else pos

    def get(self) -> Dict[str, float]:
        self.event.synchronize()
        vals = self.host.tolist()
        return {k: vals[i] for k, i in zip(self.keys, self.pos)}


class HostBatchFeeder:
    """Batches that live in (pinned) HOST memory, delivered to the device one iteration ahead of the compute stream.

    The reference keeps its whole pixel table on the GPU (train.py:14-75), so its loop never copies a batch; a caller whose table
    does not fit, or who produces batches on the host, would pay three small in-stream copies in front of every iteration
    (~0.03 ms of a 0.8 ms iteration, bench.py's `e2e` leg in round 1).  Here the copies of batch i + 1 run on a side stream
    while the kernels of batch i execute: `depth` device slots, a `ready` event per slot that the compute stream waits on, a `free`
    event per slot that the copy stream waits on before overwriting it.

        for batch in feeder.feed(host_batches):      # dicts of CPU tensors, same keys / shapes from batch to batch
            losses = trainer.step(**batch)            # device tensors, valid until the loop asks for the next batch
    """

    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        self.depth = max(2, int(depth))
        self.stream = torch.cuda.Stream(self.device)
        self._slots = [None] * self.depth
        self._ready = [torch.cuda.Event() for _ in range(self.depth)]
        self._free = [None] * self.depth
        self._head = self._tail = 0
        self.bytes_copied = 0

    def push(self, host_batch: Dict[str, torch.Tensor]) -> None:
        """Enqueues the copy of one host batch into the next free slot (side stream)."""
        if self._head - self._tail >= self.depth:
            raise RuntimeError("HostBatchFeeder: every slot holds a batch that has not been released")
        k = self._head % self.depth
        self._head += 1
        with torch.cuda.stream(self.stream):
            if self._free[k] is not None:
                self.stream.wait_event(self._free[k])  # the iteration that read this slot last has finished
            slot = self._slots[k]
            if slot is None or slot.keys() != host_batch.keys() or any(slot[n].shape != t.shape or slot[n].dtype != t.dtype for n, t in host_batch.items()):
                slot = self._slots[k] = {n: torch.empty(t.shape, dtype=t.dtype, device=self.device) for n, t in host_batch.items()}
            for n, t in host_batch.items():
                slot[n].copy_(t, non_blocking=True)
                self.bytes_copied += t.numel() * t.element_size()
            self._ready[k].record(self.stream)

    def pop(self):
        """(slot id, device batch) of the oldest pushed batch; the CURRENT stream waits for its copy, the host does not."""
        if self._tail >= self._head:
            raise RuntimeError("HostBatchFeeder: nothing pushed")
        k = self._tail % self.depth
        torch.cuda.current_stream(self.device).wait_event(self._ready[k])
        return k, self._slots[k]

    def release(self, k: int) -> None:
        """To be called once every kernel that reads slot `k` has been enqueued on the current stream."""
        ev = self._free[k] if self._free[k] is not None else torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._free[k] = ev
        self._tail += 1

    def feed(self, host_batches):
        """Generator over device batches; the copy of the next batch is enqueued before the current one is handed out."""
        it = iter(host_batches)
        nxt = next(it, None)
        if nxt is None:
            return
        self.push(nxt)
        try:
            while True:
                k, batch = self.pop()
                nxt = next(it, None)
                if nxt is not None:
                    self.push(nxt)
                yield batch
                self.release(k)
                if nxt is None:
                    return
        finally:  # the consumer left the loop early (break / exception): hand the outstanding slots back
            while self._tail < self._head:
                self.release(self._tail % self.depth)


class FusedTrainer:
    """Drop-in for the body of train()'s loop: `losses = trainer.step(xyz=..., v=..., slice_idx=...)`."""

    def __init__(self, model: NeSVoR, args: Namespace, batch_size: Optional[int] = None):
        """`batch_size`: pixels per call of `step` / `step_distributed` on THIS rank (default args.batch_size); it only
        sets the power-of-two loss scale of the fp16 backward operands."""
        self.model, self.args = model, args
        self.state = FusedState(model.inr, args, model, n_batch_samples=(batch_size or args.batch_size) * args.n_samples)
        st = self.state
        self.exp_avg = torch.zeros(st.n_train, dtype=torch.float32, device=st.device)
        self.exp_avg_sq = torch.zeros(st.n_train, dtype=torch.float32, device=st.device)
        self.lr = float(args.learning_rate)
        self.iteration = 0
        self.seed = int(getattr(args, "seed", 0) or 0)
        self.pose = not args.no_transformation_optimization
        self.dp_mode: Optional[str] = None  # set on the first distributed step: "peer" (fused kernel) or "allreduce" (NCCL / gloo)

    def _setup_dp(self, dist, world: int) -> None:
        """Chooses the data-parallel optimiser path: the fused reduce-scatter + AdamW + all-gather kernel over NVLink
        peer memory when the ranks are GPUs of one node with symmetric memory available, else all-reduce + AdamW
        (`args.dp_optimizer` = "peer" | "allreduce" forces one; a forced "peer" raises if peer memory cannot be set up)."""
        want = getattr(self.args, "dp_optimizer", "auto")
        mode = "allreduce"
        if want in ("auto", "peer") and world > 1 and self.state.device.type == "cuda" and dist.get_backend() == "nccl" and world <= 16:
            err = None
            try:
                self.state.enable_peer_memory(dist.group.WORLD)
                mode = "peer"
            except Exception as e:  # symmetric memory unavailable (no peer access, old driver, ...)
                err = e
            # the ranks must AGREE: one rank in handle.barrier() while another sits in dist.all_reduce() is a deadlock
            ok = torch.tensor([1 if mode == "peer" else 0], dtype=torch.int32, device=self.state.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                if want == "peer":
                    raise RuntimeError(f"dp_optimizer='peer': peer memory could not be set up on every rank ({err})")
                if mode == "peer":
                    self.state.disable_peer_memory()
                mode = "allreduce"
                import logging

                logging.warning("peer-memory optimiser unavailable on some rank (%s); using NCCL all-reduce + AdamW", err)
        elif want == "peer":
            raise RuntimeError("dp_optimizer='peer' needs NCCL ranks on CUDA devices (world <= 16)")
        self.dp_mode = mode

    def decay_lr(self, gamma: float) -> None:
        self.lr *= gamma

    def losses_to_host(self, losses: Dict[str, torch.Tensor]) -> LossHandle:
        """Enqueues ONE device-to-host copy of a step's loss values (pinned ring buffer) and returns its handle.  For the dict
        `step` returned (views into one slot of the loss ring) that is the copy engine alone: no gather kernel."""
        if not hasattr(self, "_host_ring"):
            self._host_ring = [torch.empty(8, dtype=torch.float32).pin_memory() for _ in range(4)]
            self._host_i = 0
        keys = list(losses.keys())
        host = self._host_ring[self._host_i % len(self._host_ring)]
        self._host_i += 1
        slot = self.state.losses
        pos = [losses[k].storage_offset() - slot.storage_offset() for k in keys]
        same = all(losses[k].untyped_storage().data_ptr() == slot.untyped_storage().data_ptr() and 0 <= i < 8 for k, i in zip(keys, pos))
        ev = torch.cuda.Event()
        cur = torch.cuda.current_stream(self.state.device)
        if same:
            # the slot stays untouched for LOSS_RING iterations, so its copy need not sit in the compute stream: a read-back
            # stream waits for the iteration's last kernel and copies while the next iteration's kernels already run
            if not hasattr(self, "_rb_stream"):
                self._rb_stream = torch.cuda.Stream(self.state.device)
                self._rb_done = [torch.cuda.Event() for _ in self._host_ring]
            done = self._rb_done[self._host_i % len(self._rb_done)]
            done.record(cur)
            with torch.cuda.stream(self._rb_stream):
                self._rb_stream.wait_event(done)
                host.copy_(slot, non_blocking=True)
                ev.record(self._rb_stream)
            self.state.last_readback = ev
        else:  # values computed elsewhere (e.g. the per-op path): gather them first
            pos = list(range(len(keys)))
            host[: len(keys)].copy_(torch.stack([losses[k].reshape(()).float() for k in keys]), non_blocking=True)
            ev.record(cur)
        h = LossHandle(keys, host, ev, pos)
        h.nbytes = 4 * (host.numel() if same else len(keys))  # what the copy moved
        return h

    def _trans_reg(self) -> None:
        """transReg (models.py:357-363) and its gradient in one native launch (`nsv_trans_reg_f32`): the weighted gradient is
        added to the flat grad's axisangle segment, the loss value lands in losses[5]."""
        st = self.state
        if not hasattr(self, "_axisangle_init"):
            self._axisangle_init = self.model.axisangle_init.detach().to(st.device, torch.float32).contiguous()
        with torch.cuda.device(st.device):
            rc = _lib.lib().nsv_trans_reg_f32(
                _lib.ptr(st.seg("axisangle")), _lib.ptr(self._axisangle_init), _lib.ptr(st.seg("axisangle", st.grad)),
                ctypes.c_void_p(st.losses.data_ptr() + 20), ctypes.c_int(self.model.n_slices),
                ctypes.c_float(float(self.args.weight_transformation)), _lib.stream(st.device))
        _lib.check(rc, "nsv_trans_reg_f32")

    def step(self, xyz, v, slice_idx, noise=None) -> Dict[str, torch.Tensor]:
        st, a = self.state, self.args
        self.iteration += 1
        st.next_losses()  # a zeroed slot of the loss ring; the parameter gradients were cleared by the previous AdamW pass
        n_q = xyz.shape[0] * a.n_samples
        losses, _ = st.forward_backward(xyz, v, slice_idx, noise, seed=self.seed, offset=(self.iteration - 1) * n_q)
        if self.pose and a.weight_transformation:
            self._trans_reg()
        out = st.loss_dict(losses)
        if self.pose and a.weight_transformation:
            out[T_REG] = losses[5]
        with torch.cuda.device(st.device):
            rc = _lib.lib().nsv_adamw_step(
                _lib.ptr(st.flat), _lib.ptr(st.grad), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), _lib.ptr(st.flat16),
                ctypes.c_int64(st.n_train), ctypes.c_float(self.lr), ctypes.c_float(0.9), ctypes.c_float(0.99), ctypes.c_float(1e-15),
                ctypes.c_float(1e-2), ctypes.c_int(self.iteration), ctypes.c_float(1.0), ctypes.c_int(1), _lib.stream(st.device))
        _lib.check(rc, "nsv_adamw_step")
        return out

    def step_distributed(self, dist, world: int, xyz, v, slice_idx, noise=None) -> Dict[str, torch.Tensor]:
        """Data-parallel iteration (one process per GPU): every rank runs kernel A on its own B pixels,
        the flat gradient of the trainable prefix is summed with ONE NCCL all-reduce, and the mean over
        ranks is folded into AdamW's unscale factor, so all replicas apply the identical update."""
        st, a = self.state, self.args
        if self.dp_mode is None:
            self._setup_dp(dist, world)
        self.iteration += 1
        st.next_losses()
        n_q = xyz.shape[0] * a.n_samples
        rank = dist.get_rank()
        losses, _ = st.forward_backward(xyz, v, slice_idx, noise, seed=self.seed + 7919 * rank,
                                        offset=(self.iteration - 1) * n_q, dist=dist, world=world)
        if self.pose and a.weight_transformation:
            self._trans_reg()  # identical on every rank: the all-reduce mean leaves it unchanged
        out = st.loss_dict(losses)
        if self.pose and a.weight_transformation:
            out[T_REG] = losses[5]
        self._dp_update(dist, world)
        return out

    def _dp_update(self, dist, world: int) -> None:
        """Mean of the ranks' gradients + AdamW + refreshed fp16 parameters on every rank; clears the gradient."""
        st = self.state
        rank = dist.get_rank()
        if self.dp_mode is None:
            self._setup_dp(dist, world)
        if self.dp_mode == "peer" and getattr(self.args, "dp_sync", None) is None:
            import os

            # "host": two symmetric-memory barriers around the kernel (default: measured faster at 8 ranks, 0.863 vs 0.896 ms per
            # step, profiles/r02_bench_8gpu_weak*.json -- every block of the in-kernel variant pays a system-scope fence for
            # its peer writes); "kernel": the ranks' rendezvous inside the kernel (nsv_adamw_step_dp_sync), equal at 2-4 ranks
            self.args.dp_sync = os.environ.get("NSV_DP_SYNC", "host")
        if self.dp_mode == "peer" and self.args.dp_sync in ("kernel", "hybrid"):
            # "kernel": ONE launch -- rendezvous of the ranks (flags in peer memory), reduce-scatter, AdamW, all-gather, rendezvous.
            # "hybrid": the rendezvous BEFORE inside the kernel (flags only: kernel A has completed, nothing to fence), the one
            # AFTER as a host-launched symmetric-memory barrier (the in-kernel one costs every block a system-scope fence).
            mode = 3 if self.args.dp_sync == "kernel" else 1
            with torch.cuda.device(st.device):
                rc = _lib.lib().nsv_adamw_step_dp_sync(
                    _lib.ptr(st.flat), st.peer_grads, _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), st.peer_flat16,
                    ctypes.c_int(world), ctypes.c_int(rank), ctypes.c_int64(st.n_train), ctypes.c_float(self.lr), ctypes.c_float(0.9),
                    ctypes.c_float(0.99), ctypes.c_float(1e-15), ctypes.c_float(1e-2), ctypes.c_int(self.iteration),
                    ctypes.c_float(1.0 / world), ctypes.c_int64(st.tail_lo), st.peer_tail32, st.peer_flags,
                    ctypes.c_uint64(self.iteration), ctypes.c_int(mode), _lib.stream(st.device))
            _lib.check(rc, "nsv_adamw_step_dp_sync")
            if mode == 1:
                st.peer_handle.barrier()  # every owner has read this rank's gradient and written this rank's copies
            st.grad[: st.n_train].zero_()
            return
        if self.dp_mode == "peer":
            if getattr(self.args, "dp_multimem", None) is None:
                import os

                # unicast peer loads / stores or the NVSwitch-multicast variant (nsv_adamw_step_dp_mc)?  Measured, not guessed:
                # 2 ranks 0.086 vs 0.122 ms, 4 ranks 0.110 vs 0.111 ms (profiles/r02_dp_multimem_*.json) -- unicast traffic grows with
                # the number of peers, multicast traffic does not, so the answer depends on the world size and the fabric.
                # NSV_DP_MULTIMEM=0 / 1 forces one; otherwise a fresh trainer times both on its own box before its first update.
                env = os.environ.get("NSV_DP_MULTIMEM", "auto")
                if env in ("0", "1"):
                    self.args.dp_multimem = env == "1"
                elif st.mc_ptrs is not None and self.iteration <= 1:
                    self.args.dp_multimem = self._autotune_multimem(dist, world, rank)
                else:
                    self.args.dp_multimem = False
            mc = st.mc_ptrs if self.args.dp_multimem else None
            self._peer_exchange(world, rank, mc, self.lr, 1.0 / world, self.iteration)
            self.dp_multimem_active = mc is not None
            st.grad[: st.n_train].zero_()
            return
        from .distributed import allreduce_gradient

        unscale = allreduce_gradient(st.grad[: st.n_train], dist, world)
        with torch.cuda.device(st.device):
            rc = _lib.lib().nsv_adamw_step(
                _lib.ptr(st.flat), _lib.ptr(st.grad), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), _lib.ptr(st.flat16),
                ctypes.c_int64(st.n_train), ctypes.c_float(self.lr), ctypes.c_float(0.9), ctypes.c_float(0.99), ctypes.c_float(1e-15),
                ctypes.c_float(1e-2), ctypes.c_int(self.iteration), ctypes.c_float(unscale), ctypes.c_int(1), _lib.stream(st.device))
        _lib.check(rc, "nsv_adamw_step")

    def _peer_exchange(self, world: int, rank: int, mc, lr: float, unscale: float, step: int) -> None:
        """barrier | reduce-scatter + AdamW + all-gather in one kernel over peer memory (unicast, or multicast when `mc`) | barrier."""
        st = self.state
        h = st.peer_handle
        h.barrier()  # every rank's kernel A has finished: all gradients are complete
        common = (_lib.ptr(st.flat), st.peer_grads, _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), st.peer_flat16,
                  ctypes.c_int(world), ctypes.c_int(rank), ctypes.c_int64(st.n_train), ctypes.c_float(lr), ctypes.c_float(0.9),
                  ctypes.c_float(0.99), ctypes.c_float(1e-15), ctypes.c_float(1e-2), ctypes.c_int(step),
                  ctypes.c_float(unscale), ctypes.c_int64(st.tail_lo), st.peer_tail32)
        with torch.cuda.device(st.device):
            if mc is not None:
                rc = _lib.lib().nsv_adamw_step_dp_mc(*common, ctypes.c_void_p(mc[0]), ctypes.c_void_p(mc[1]), ctypes.c_void_p(mc[2] or None),
                                                     _lib.stream(st.device))
            else:
                rc = _lib.lib().nsv_adamw_step_dp(*common, _lib.stream(st.device))
        _lib.check(rc, "nsv_adamw_step_dp_mc" if mc is not None else "nsv_adamw_step_dp")
        h.barrier()  # every owner has read this rank's gradient and written this rank's fp16 parameters

    def _autotune_multimem(self, dist, world: int, rank: int) -> bool:
        """Times the unicast and the multicast exchange on this box (a few launches each, before the first real update) and returns
        True when multicast is faster on the slowest rank.  The timed launches run with lr = 0 and a zero gradient scale on a trainer
        whose moments are still zero: parameters, moments and the fp16 / fp32 copies come out bit-identical, the pending gradient is
        only read."""
        st = self.state
        times = []
        for mc in (None, st.mc_ptrs):
            for _ in range(2):
                self._peer_exchange(world, rank, mc, 0.0, 0.0, 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(8):
                self._peer_exchange(world, rank, mc, 0.0, 0.0, 1)
            e1.record()
            torch.cuda.synchronize(st.device)
            times.append(e0.elapsed_time(e1) / 8)
        t = torch.tensor(times, dtype=torch.float32, device=st.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the same two numbers on every rank: the same choice on every rank
        uni, mcast = (float(x) for x in t.tolist())
        self.dp_autotune_ms = {"unicast": uni, "multimem": mcast}
        return mcast < 0.97 * uni  # ties go to the unicast kernel (bit-identical to all-reduce + AdamW in rank order)

    def sync_to_model(self) -> None:
        if self.dp_mode == "peer" and getattr(self.state, "dp_flags", None) is not None and int(self.state.dp_flags[33]) != 0:
            raise RuntimeError(f"data-parallel optimiser: a rendezvous timed out at step {int(self.state.dp_flags[33])} (a rank stopped responding)")
        if self.dp_mode == "peer":  # every rank holds only its own shard of the fp32 master: collect the others
            import torch.distributed as dist

            st, world = self.state, dist.get_world_size()
            lo, hi = ctypes.c_int64(), ctypes.c_int64()
            for r in range(world):
                _lib.lib().nsv_adamw_shard_bounds(ctypes.c_int64(st.n_train), ctypes.c_int(world), ctypes.c_int(r), ctypes.byref(lo), ctypes.byref(hi))
                if hi.value > lo.value:
                    dist.broadcast(st.flat[lo.value : hi.value], src=r)
        self.state.push_to_model()


def fused_render(inr: INR, xyz: torch.Tensor, transformation: Optional[RigidTransform], psf_sigma, n_samples: int,
                 noise: Optional[torch.Tensor] = None, seed: int = 0, state: Optional[FusedState] = None) -> torch.Tensor:
    """INR.sample_batch + INR.forward(...).mean(-1) (sample.py:25-31,44-50) in one launch."""
    if state is None:
        state = getattr(inr, "_fused_state", None)
        if state is None:
            raise RuntimeError("fused_render: attach a FusedState first (attach_render_state(inr, args))")
    M = xyz.shape[0]
    S = max(int(n_samples), 1)
    xyz = xyz.contiguous().float()
    mat = None
    per_point = 0
    if transformation is not None:
        mat = transformation.matrix(True).contiguous().float()
        per_point = int(mat.shape[0] > 1)
        assert mat.shape[0] in (1, M)
    if isinstance(psf_sigma, torch.Tensor):
        sig = psf_sigma.to(xyz.device).float().reshape(-1, 3).contiguous() if psf_sigma.numel() > 1 else psf_sigma.to(xyz.device).float().reshape(1).expand(3).contiguous()
    elif isinstance(psf_sigma, (tuple, list)):
        sig = torch.tensor([list(psf_sigma)], dtype=torch.float32, device=xyz.device)
    else:
        sig = torch.full((1, 3), float(psf_sigma), dtype=torch.float32, device=xyz.device)
    sig_pp = int(sig.numel() > 3)
    out = torch.empty(M, dtype=torch.float32, device=xyz.device)
    prm = state.params_struct()
    if noise is not None:
        noise = noise.contiguous().float()
    # every launch continues the Philox stream where the previous one stopped: the reference draws fresh randn per batch
    # (models.py:161-169), so the Monte-Carlo error of successive inference chunks / slices must be independent
    offset = getattr(state, "render_offset", 0)
    state.render_offset = offset + M * S
    with torch.cuda.device(xyz.device):
        rc = _lib.lib().nsv_inr_render(
            ctypes.byref(state.cfg), ctypes.byref(prm), _lib.ptr(xyz), _lib.ptr(mat), ctypes.c_int(per_point), _lib.ptr(sig),
            ctypes.c_int(sig_pp), _lib.ptr(noise), ctypes.c_uint64(seed), ctypes.c_uint64(offset), _lib.ptr(out), ctypes.c_int64(M),
            ctypes.c_int(S), _lib.stream(xyz.device))
    if rc == -2:
        raise FusedUnsupported(_lib.lib().nsv_last_error_string().decode())
    _lib.check(rc, "nsv_inr_render")
    return out


def attach_render_state(inr: INR, args: Namespace) -> FusedState:
    state = FusedState(inr, args, None)
    inr._fused_state = state
    return state
