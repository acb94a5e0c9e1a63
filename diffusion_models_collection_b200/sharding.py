"""Multi-GPU sampling: the sample batch is partitioned across the ranks of one node (one process per GPU).

Samples are independent everywhere on the path (no cross-sample op in UNet/DiT forward, GroupNorm and the dynamic
threshold quantile are per sample -- SURVEY.md section 8e), so the 50-1000 denoising steps need NO exchange: rank r
denoises the contiguous slice [lo_r, hi_r) of the batch and of the labels, using the r-th slice of the SAME global
x_T noise tensor (so an N-GPU run reproduces the 1-GPU run image for image), and only the final images are
all-gathered (NCCL over NVLink / NVSwitch; gloo on CPU in the tests).  The reference has no multi-GPU sampling
(single-device loop, sample.py:180-206); this is the capability BASELINE.json adds.
"""

from __future__ import annotations

import contextlib

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int):
    """[lo, hi) of rank's contiguous slice; the first (total % world) ranks get one extra sample."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, rem = divmod(int(total), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_shards(local: torch.Tensor, total: int, rank: int, world: int, group=None) -> torch.Tensor:
    """all-gather of per-rank slices (possibly uneven, possibly empty) along dim 0 -> [total, ...] on every rank."""
    if world == 1:
        return local
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    tail = tuple(local.shape[1:])
    pad = local.new_zeros((mx,) + tail)
    pad[: local.shape[0]] = local
    out = local.new_empty((world * mx,) + tail)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if all(hi - lo == mx for lo, hi in sizes):
        return out
    return torch.cat([out[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)


def sharded_call(fn, shape, y=None, noise=None, rank=None, world=None, sliced=False, group=None, diffusion=None):
    """Runs fn(local_shape, y_local, noise_local) -> [n_local, ...] on this rank's slice and gathers the result.
    y / noise are GLOBAL tensors (sliced here) unless sliced=True."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    total = int(shape[0])
    lo, hi = shard_bounds(total, rank, world)
    if not sliced:
        y = None if y is None else y[lo:hi]
        noise = None if noise is None else noise[lo:hi]
    local_shape = (hi - lo,) + tuple(shape[1:])
    if hi > lo:
        with _noise_shard(diffusion, (total, lo, hi) if world > 1 else None):
            local = fn(local_shape, y, noise)
    else:  # more ranks than samples: nothing to do here, but still take part in the gather
        ref = noise if noise is not None else torch.empty(0)
        local = torch.empty((0,) + tuple(shape[1:]), dtype=torch.float32, device=ref.device)
    return gather_shards(local, total, rank, world, group)


def _resolve(rank, world, group):
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    return rank, world


@contextlib.contextmanager
def _noise_shard(diffusion, shard):
    """per-step draws (DDPM, DDIM eta > 0) of a shard = this rank's rows of the GLOBAL per-step draw (seed parity with the
    single-process run); a no-op for samplers without the hook"""
    if diffusion is None:
        yield
        return
    prev = getattr(diffusion, "_noise_shard", None)
    if hasattr(diffusion, "_draw_like"):
        diffusion._noise_shard = shard
    try:
        yield
    finally:
        if hasattr(diffusion, "_draw_like"):
            diffusion._noise_shard = prev


def _global_noise(diffusion, shape, noise, rank, world, sliced):
    """x_T: the caller's tensor, else ONE global draw from the generator (same stream position on every rank, so an
    N-rank run reproduces the 1-rank run), cut to this rank's slice when the other inputs are already local"""
    if noise is not None:
        return noise
    noise = torch.randn(shape, device=diffusion.device)
    if sliced:
        lo, hi = shard_bounds(int(shape[0]), rank, world)
        noise = noise[lo:hi]
    return noise


def sharded_sample_with_cfg(diffusion, model, shape, y, cfg_scale=3.0, p_threshold=0.995, noise=None, rank=None,
                            world=None, sliced=False, group=None):
    """diffusion.sample_with_cfg over this rank's slice of the batch + all-gather of the final images.
    y / noise: global [B, ...] tensors, or this rank's slices when sliced=True."""
    rank, world = _resolve(rank, world, group)
    noise = _global_noise(diffusion, shape, noise, rank, world, sliced)

    def fn(local_shape, yl, nl):
        return diffusion.sample_with_cfg(model, local_shape, yl, cfg_scale=cfg_scale, p_threshold=p_threshold, noise=nl)

    return sharded_call(fn, shape, y, noise, rank, world, sliced, group, diffusion=diffusion)


def sharded_sample(diffusion, model, shape, y=None, noise=None, rank=None, world=None, sliced=False, group=None):
    """diffusion.sample over this rank's slice of the batch + all-gather of the final images."""
    rank, world = _resolve(rank, world, group)
    noise = _global_noise(diffusion, shape, noise, rank, world, sliced)

    def fn(local_shape, yl, nl):
        return diffusion.sample(model, local_shape, yl, noise=nl)

    return sharded_call(fn, shape, y, noise, rank, world, sliced, group, diffusion=diffusion)
