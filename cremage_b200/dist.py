"""Multi-GPU plumbing of the hot path: one process per GPU, the batch of independent images is partitioned across
ranks (no collective inside the UNet step), and the decoded uint8 images are gathered once per job.

The reference is single-GPU (no parallelism anywhere on the path); seeds/prompts are independent
units, so this is plain sharding.  For bit-parity with a one-GPU run seeded by a single `randn(B, ...)`
(k_diffusion_samplers.py:169), every rank draws the FULL-batch noise with the same seed and keeps its slice.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous rows [lo, hi) of rank `rank`; the first `global_batch % world` ranks take one extra row."""
    if global_batch < 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad shard request: batch={global_batch} rank={rank} world={world}")
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors: Sequence[torch.Tensor], rank: int, world: int):
    """Slice every tensor's leading (batch) dimension to this rank's rows."""
    out = []
    for t in tensors:
        lo, hi = shard_range(t.shape[0], rank, world)
        out.append(t[lo:hi])
    return out


def full_batch_noise(shape: Sequence[int], seed: int, rank: int, world: int, device="cpu", dtype=torch.float32):
    """randn(shape) drawn identically on every rank (CPU generator, same seed), sliced to this rank's rows."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    full = torch.randn(tuple(shape), generator=g, dtype=dtype)
    lo, hi = shard_range(shape[0], rank, world)
    return full[lo:hi].to(device)


def gather_images(local: torch.Tensor, global_batch: int) -> torch.Tensor:
    """all_gather of decoded images [b_local, H, W, 3] -> [global_batch, H, W, 3] in rank order (ragged shards are
    padded to the largest shard for the collective and trimmed afterwards). No-op without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(global_batch, r, world) for r in range(world)]
    bmax = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < bmax:
        pad = torch.cat([local, local.new_zeros((bmax - local.shape[0], *local.shape[1:]))])
    out = local.new_empty((world * bmax, *local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous())
    parts = [out[r * bmax: r * bmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts)
