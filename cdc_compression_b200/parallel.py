"""Batch sharding of the decode across GPUs (one process per GPU, ``torch.distributed``).

The path shards naturally: no operator mixes batch elements.  Rank r decodes images
[lo_r, hi_r) — running ``context_fn`` on its own shard, so no context tensor crosses NVLink — and the
only collective is one all-gather of the decoded images at the end.  The init noise is drawn for the
WHOLE batch before the split, so a G-GPU decode is bit-identical per image to the 1-GPU decode — for the
deterministic sampler (eta == 0) and the per-image clip modes ("none" / "full").  Two settings couple the result to
the shard and are rejected when world_size > 1: eps ``clip_noise="half"`` (the reference clamps the first B/2 images
of the batch it is given, denoising_diffusion.py:142-143 — of the LOCAL shard here) and ``eta != 0`` (every rank
would draw its own noise stream).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n items; the first n % world ranks take one extra item."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_batch(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather row-shards produced with ``shard_range`` back into the full batch (uneven shards ok)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


def sharded_decode(decode_fn: Callable[..., Tuple[torch.Tensor, torch.Tensor]], images: torch.Tensor,
                   init: Optional[torch.Tensor] = None, group=None, **kwargs):
    """Run ``decode_fn(images_shard, init=init_shard, **kwargs) -> (x_hat, bpp_per_image)`` on this rank's
    shard of the batch and gather both results.  ``decode_fn`` is typically
    ``functools.partial(diffusion.compress, sample_steps=S, bpp_return_mean=False, ...)``."""
    n = images.shape[0]
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    if world > 1:
        owner = getattr(getattr(decode_fn, "func", decode_fn), "__self__", None)   # functools.partial(d.compress, ...)
        eta = kwargs.get("eta", getattr(decode_fn, "keywords", {}).get("eta", 0))
        if eta:
            raise NotImplementedError("sharded_decode: eta != 0 draws per-rank noise streams; decode on one rank or "
                                      "pass eta=0")
        if getattr(owner, "clip_noise", None) == "half":
            raise NotImplementedError("sharded_decode: clip_noise='half' clamps the first half of the LOCAL batch; "
                                      "use 'none' or 'full' when sharding")
    lo, hi = shard_range(n, rank, world)
    if hi > lo:
        x_hat, bpp = decode_fn(images[lo:hi], init=None if init is None else init[lo:hi], **kwargs)
        bpp = bpp.reshape(-1)
    else:  # more ranks than images
        x_hat = images.new_zeros((0,) + tuple(images.shape[1:]))
        bpp = images.new_zeros((0,))
    return gather_batch(x_hat, n, group), gather_batch(bpp, n, group)
