"""Data-parallel sharding of the image batch over the GPUs of one box.

The guided loop has no cross-image coupling (per-sample guidance), so each rank runs its contiguous
shard independently - no collective inside the loop - and the only communication is one all_gather
of the final images (NCCL on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of ``n`` samples for ``rank``; the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sample_seeds(base_seed: int, lo: int, hi: int) -> List[int]:
    """Per-SAMPLE seeds, so results do not depend on the number of ranks."""
    return [base_seed + i for i in range(lo, hi)]


def gather_images(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All ranks' shards concatenated in rank order -> (n_total, ...).  Ragged shards are padded to
    the largest shard for the collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(n_total, world, r) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    buf = local
    if local.shape[0] < biggest:
        pad = torch.zeros((biggest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf = torch.cat([local, pad])
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf.contiguous())
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)])
