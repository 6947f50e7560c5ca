"""Clip-sharded multi-GPU extraction: one process per GPU, contiguous clip ranges per rank, no
collective on the compute path; ONE all-gather of the rank's (B/G, M, T) mel block when every rank
needs all frames.  The reference has no distributed code (only nn.DataParallel key
stripping, trainer.py:269-272), so this is new surface, not a mirror."""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_clips: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced clip range [start, stop) of `rank` (first n % G ranks get one extra clip)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(n_clips, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_gather_mel(local: torch.Tensor, n_clips: Optional[int] = None, group=None) -> torch.Tensor:
    """Gather per-rank (b_r, M, T) blocks into (sum b_r, M, T) on every rank.

    Equal shards use a single `all_gather_into_tensor` (NCCL over NVLink on GPUs; dim-0 contiguous so no
    re-layout).  Ragged shards (n_clips not divisible by the world size) pad the short ranks by one clip
    for the collective and drop the padding afterwards."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return local
    b_local = local.shape[0]
    if n_clips is None:
        n_clips = b_local * world
    sizes = [shard_range(n_clips, r, world) for r in range(world)]
    counts = [b - a for a, b in sizes]
    if counts[rank] != b_local:
        raise ValueError(f"rank {rank} holds {b_local} clips, expected {counts[rank]}")
    bmax = max(counts)
    local = local.contiguous()
    if bmax != b_local:
        padded = local.new_zeros((bmax,) + tuple(local.shape[1:]))
        padded[:b_local] = local
        local = padded
    out = local.new_empty((world * bmax,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local, group=group)
    if all(c == bmax for c in counts):
        return out
    return torch.cat([out[r * bmax:r * bmax + counts[r]] for r in range(world)], dim=0)


class ShardedExtractor:
    """Run `module` on this rank's contiguous shard of a clip batch that every rank holds (or can
    index), optionally gathering the result.  `module(wav (b, L)) -> (b, M, T)`."""

    def __init__(self, module, group=None):
        self.module = module
        self.group = group

    def __call__(self, wav_all: torch.Tensor, gather: bool = True) -> torch.Tensor:
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        a, b = shard_range(wav_all.shape[0], rank, world)
        local = self.module(wav_all[a:b])
        if gather and world > 1:
            return all_gather_mel(local, wav_all.shape[0], self.group)
        return local
