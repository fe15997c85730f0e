"""Training driver of the INR path: pixel table, optimiser, hot loop.

Host-side mirror of nesvor/nesvor/train.py: `Dataset` (:14-120: pixel table, `bounding_box`, `mean`,
`get_batch` with full-table reshuffle at epoch end, `mask`) and `train(slices, args)` (:123-232:
AdamW with two groups, MultiStepLR on milestones, GradScaler(init_scale=1), loss weights, logging
of moving averages, outputs).  The per-iteration compute is delegated to `NeSVoR.forward`
(autograd-composed native ops under fp16 autocast, like the reference's loop) or, when `args.fused` is set, to the fused
kernel + fused AdamW (`fused.FusedTrainer`); a configuration the fused kernels are not instantiated for falls back to the
per-op path with a warning (single process; under data parallelism it raises).

Data parallelism (the reference is single-GPU): when `torch.distributed` is initialised with more than
one rank, `train` keeps the pixel table replicated, makes every rank draw the SAME global batch of `args.batch_size`
pixels (the epoch permutation is broadcast from rank 0), hands rank r the r-th contiguous chunk of it and steps with
`FusedTrainer.step_distributed` (gradient mean + AdamW over NVLink peer memory or one NCCL all-reduce).
"""
from argparse import Namespace
from typing import Dict, List, Tuple
import datetime
import logging
import time

import torch
import torch.optim as optim

from ..image import Slice, Volume
from ..transform import RigidTransform, transform_points
from ..utils import MovingAverage, gaussian_blur
from .models import B_REG, D_LOSS, DS_LOSS, I_REG, INR, S_LOSS, T_REG, NeSVoR


class Dataset(object):
    def __init__(self, slices: List[Slice], args: Namespace) -> None:
        self.mask_threshold = getattr(args, "mask_threshold", 1.0)
        xyz_all, v_all, slice_idx_all, transformation_all, resolution_all = [], [], [], [], []
        for i, s in enumerate(slices):
            v = s.v_masked
            xyz_all.append(s.xyz_masked_untransformed)
            v_all.append(v)
            slice_idx_all.append(torch.full(v.shape, i, device=v.device))
            transformation_all.append(s.transformation)
            resolution_all.append(s.resolution_xyz)
        self.xyz = torch.cat(xyz_all)
        self.v = torch.cat(v_all)
        self.slice_idx = torch.cat(slice_idx_all)
        self.transformation = RigidTransform.cat(transformation_all)
        self.resolution = torch.stack(resolution_all, 0)
        self.count = self.v.shape[0]
        self.epoch = 0
        self.dist = None  # set by train() under data parallelism: the epoch permutation is then broadcast from rank 0
        # locality-aware batch order (SURVEY s.8f row 1): after the epoch shuffle the pixels of EVERY batch are put in
        # (slice, row, column) order -- the batches keep the reference's composition (train.py:60-75: consecutive blocks of
        # one randperm; every loss is a mean over the batch, so the order inside a batch is immaterial to the optimisation),
        # but consecutive tiles of kernel A now belong to neighbouring pixels of one slice and share coarse / mid-level
        # table entries in L1 / L2.  One sort per epoch (P keys), nothing per iteration.
        self.locality_batch = int(getattr(args, "locality_batch_size", 0) or 0)

    @property
    def xyz_transformed(self) -> torch.Tensor:
        return transform_points(self.transformation[self.slice_idx], self.xyz)

    @property
    def bounding_box(self) -> torch.Tensor:
        max_r = self.resolution.max()
        xyz = self.xyz_transformed
        return torch.stack([xyz.amin(0) - 2 * max_r, xyz.amax(0) + 2 * max_r], 0)

    @property
    def mean(self) -> float:
        v = self.v if self.v.numel() < 256**3 else self.v[: 256**3]
        q1, q2 = torch.quantile(v, torch.tensor([0.1, 0.9], dtype=v.dtype, device=v.device))
        return self.v[torch.logical_and(self.v > q1, self.v < q2)].mean().item()

    def get_batch(self, batch_size: int, device) -> Dict[str, torch.Tensor]:
        if self.count + batch_size > self.xyz.shape[0]:  # new epoch: shuffle the whole table
            self.count = 0
            self.epoch += 1
            idx = torch.randperm(self.xyz.shape[0], device=device)
            if self.dist is not None:
                self.dist.broadcast(idx, src=0)
            if getattr(self, "locality_batch", 0):
                idx = self._order_inside_batches(idx, self.locality_batch)
            self.xyz, self.v, self.slice_idx = self.xyz[idx], self.v[idx], self.slice_idx[idx]
        sl = slice(self.count, self.count + batch_size)
        self.count += batch_size
        return {"xyz": self.xyz[sl], "v": self.v[sl], "slice_idx": self.slice_idx[sl]}

    def _order_inside_batches(self, idx: torch.Tensor, batch_size: int) -> torch.Tensor:
        """`idx` = the epoch's permutation of the pixel table; returns it with every consecutive block of `batch_size`
        entries sorted by (slice, y, x) of the pixels they select.  Keys are built so that ONE stable sort does it."""
        n = idx.numel()
        xyz = self.xyz[idx]
        lo = self.xyz.amin(0)
        step = float(self.resolution[:, :2].min())
        ix = ((xyz[:, 0] - lo[0]) / step).round().long().clamp_(0, (1 << 13) - 1)
        iy = ((xyz[:, 1] - lo[1]) / step).round().long().clamp_(0, (1 << 13) - 1)
        block = torch.arange(n, device=idx.device) // batch_size
        key = ((block << 20 | self.slice_idx[idx].long()) << 26) | (iy << 13) | ix
        return idx[torch.argsort(key)]

    @property
    def mask(self) -> Volume:
        """Occupancy mask of the transformed pixel cloud: bincount + separable blur (train.py:82-120)."""
        with torch.no_grad():
            r_min, r_max = self.resolution.min(), self.resolution.max()
            xyz = self.xyz_transformed
            xyz_min = xyz.amin(0) - r_max * 10
            xyz_max = xyz.amax(0) + r_max * 10
            shape_xyz = ((xyz_max - xyz_min) / r_min).ceil().long()
            shape = (int(shape_xyz[2]), int(shape_xyz[1]), int(shape_xyz[0]))
            kji = ((xyz - xyz_min) / r_min).round().long()
            flat = kji[..., 0] + shape[2] * kji[..., 1] + shape[2] * shape[1] * kji[..., 2]
            mask = torch.bincount(flat, minlength=shape[0] * shape[1] * shape[2]).view((1, 1) + shape).float()
            thr = self.mask_threshold * r_min**3 / self.resolution.log().mean().exp() ** 3
            thr = thr * (mask.sum() / (mask > 0).sum())
            mask = (gaussian_blur(mask, (r_max / r_min).item(), 3) > thr)[0, 0]
            xyz_c = xyz_min + (shape_xyz - 1) / 2 * r_min
            return Volume(mask.float(), mask, RigidTransform(torch.cat([0 * xyz_c, xyz_c])[None], True), r_min, r_min, r_min)


def build_optimizer(model: torch.nn.Module, args: Namespace):
    """AdamW with the reference's two parameter groups (train.py:134-152)."""
    params_net, params_encoding = [], []
    for name, param in model.named_parameters():
        if param.numel() > 0:
            (params_net if "_net" in name else params_encoding).append(param)
    return torch.optim.AdamW(
        params=[{"name": "encoding", "params": params_encoding}, {"name": "net", "params": params_net, "weight_decay": 1e-2}],
        lr=args.learning_rate, betas=(0.9, 0.99), eps=1e-15)


def loss_weights(args: Namespace) -> Dict[str, float]:
    return {D_LOSS: 1, S_LOSS: 1, T_REG: args.weight_transformation, B_REG: args.weight_bias, I_REG: args.weight_image}


def train(slices: List[Slice], args: Namespace) -> Tuple[INR, List[Slice], Volume]:
    dataset = Dataset(slices, args)
    model = NeSVoR(dataset.transformation, dataset.resolution, dataset.mean, dataset.bounding_box, args)
    use_fused = bool(getattr(args, "fused", False))
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    if world > 1:
        from .distributed import shard_batch

        if not use_fused:
            raise RuntimeError("data-parallel training runs on the fused path: set args.fused")
        if args.batch_size % world:
            raise RuntimeError(f"args.batch_size ({args.batch_size}) must be a multiple of the number of ranks ({world})")
        dataset.dist = dist
        with torch.no_grad():  # replicas start from rank 0's initialisation (nn.Embedding draws from the global RNG)
            for p in model.parameters():
                dist.broadcast(p.data, src=0)
    first_losses = None
    if use_fused:
        from .fused import FusedTrainer, FusedUnsupported

        try:  # the constructor checks the configuration, the first launch checks the kernel instantiations
            trainer = FusedTrainer(model, args, batch_size=args.batch_size // world)
            batch = dataset.get_batch(args.batch_size, args.device)
            first_losses = (trainer.step_distributed(dist, world, **shard_batch(batch, rank, world)) if world > 1 else trainer.step(**batch))
        except FusedUnsupported as e:
            if world > 1:
                raise
            logging.warning("nesvor_b200: %s -- training on the per-op native path", e)
            use_fused = False
    if not use_fused:
        optimizer = build_optimizer(model, args)
        scheduler = optim.lr_scheduler.MultiStepLR(optimizer=optimizer, milestones=list(range(1, len(args.milestones) + 1)), gamma=args.gamma)
        fp16 = not args.single_precision
        scaler = torch.amp.GradScaler("cuda", init_scale=1.0, enabled=fp16, growth_factor=2.0, backoff_factor=0.5)
    decay_milestones = [int(m * args.n_iter) for m in args.milestones]
    model.train()
    weights = loss_weights(args)
    average = MovingAverage(1 - 0.001)
    logging.info("NeSVoR training starts.")
    train_time = 0.0
    pending = None
    for i in range(1, args.n_iter + 1):
        t0 = time.time()
        if first_losses is not None:  # iteration 1 ran above (it doubles as the probe of the fused instantiations)
            losses, first_losses = first_losses, None
        elif world > 1:
            losses = trainer.step_distributed(dist, world, **shard_batch(dataset.get_batch(args.batch_size, args.device), rank, world))
        elif use_fused:
            losses = trainer.step(**dataset.get_batch(args.batch_size, args.device))
        else:
            batch = dataset.get_batch(args.batch_size, args.device)
            with torch.autocast("cuda", dtype=torch.float16, enabled=fp16):  # train.py:183: forward and loss sum under autocast
                losses = model(**batch)
                loss = 0
                for k in losses:
                    if k in weights and weights[k]:
                        loss = loss + weights[k] * losses[k]
            scaler.scale(loss).backward()
            if getattr(args, "debug", False):
                for _name, _p in model.named_parameters():
                    if _p.grad is not None and not _p.grad.isfinite().all():
                        logging.debug("iter %d: Found NaNs in the grad of %s", i, _name)
            scaler.step(optimizer)
            scaler.update()
            optimizer.zero_grad()
        train_time += time.time() - t0
        if not getattr(args, "no_loss_sync", False):
            if use_fused:  # one async D2H copy per step, read one step late: the GPU never drains (LossHandle)
                if pending is not None:
                    for k, val in pending.get().items():
                        average(k, val)
                pending = trainer.losses_to_host(losses)
                if i == args.n_iter or (decay_milestones and i >= decay_milestones[0]):
                    for k, val in pending.get().items():  # a log line is due: catch up
                        average(k, val)
                    pending = None
            else:
                for k in losses:
                    average(k, losses[k].item())
        if (decay_milestones and i >= decay_milestones[0]) or i == args.n_iter:
            lr = trainer.lr if use_fused else optimizer.param_groups[0]["lr"]
            logging.info("time %s epoch %d iter %d %s lr %.3e", datetime.timedelta(seconds=int(train_time)), dataset.epoch, i,
                         " ".join("%s=%.4e" % (k, average[k]) for k in losses), lr)
            if i < args.n_iter:
                decay_milestones.pop(0)
                if use_fused:
                    trainer.decay_lr(args.gamma)
                else:
                    scheduler.step()
    if use_fused:
        trainer.sync_to_model()
    transformation = model.transformation
    dataset.transformation = transformation
    mask = dataset.mask
    output_slices = []
    for i in range(len(slices)):
        output_slice = slices[i].clone()
        output_slice.transformation = transformation[i]
        output_slices.append(output_slice)
    return model.inr, output_slices, mask
