"""Multi-GPU host logic: how a batch of independent frames is split over ranks and how the per-rank
results are stitched back into one stream index.

Frames never depend on each other, so there is no data-path collective: every rank compresses /
decompresses its own contiguous range of chunks.  The only exchange is an all_gather of the per-frame
sizes (8 bytes per frame) from which every rank derives the global offsets of the concatenated stream.
Works on any torch.distributed backend (NCCL on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first (n % world) ranks get one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_sizes(local_sizes: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """all_gather the per-frame sizes of every rank -> tensor of n_items sizes in global frame order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_sizes.clone()
    counts = [shard_range(n_items, r, world) for r in range(world)]
    longest = max(e - s for s, e in counts)
    padded = torch.zeros(longest, dtype=local_sizes.dtype, device=local_sizes.device)
    padded[: local_sizes.numel()] = local_sizes
    bufs = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    return torch.cat([bufs[r][: counts[r][1] - counts[r][0]] for r in range(world)])


def global_offsets(all_sizes: torch.Tensor) -> torch.Tensor:
    """Exclusive scan: offsets[i] = start of frame i in the concatenated stream; offsets[n] = total."""
    out = torch.zeros(all_sizes.numel() + 1, dtype=torch.int64, device=all_sizes.device)
    out[1:] = torch.cumsum(all_sizes.to(torch.int64), 0)
    return out


def rank_byte_range(offsets: torch.Tensor, n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Where rank's frames live in the concatenated stream (for gatherv-style placement)."""
    s, e = shard_range(n_items, rank, world)
    return int(offsets[s]), int(offsets[e])
