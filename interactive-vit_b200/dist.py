"""Data-parallel sharding of image batches over the GPUs of one box (one process per GPU).

The path shards over independent units — images: weights are replicated, each rank runs the
whole forward on a contiguous slice of the batch, and the only exchange step is a gather of the (small) results
to rank 0.  The reference has nothing here (single process, CPU); `torch.distributed` is plumbing: NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, start+count) slice of `total` images owned by `rank`; the first `total % world` ranks
    get one extra image, so counts differ by at most one and empty shards only occur when total < world."""
    if total < 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad shard request total={total} world={world} rank={rank}")
    base, extra = divmod(total, world)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def gather_to_rank0(local: torch.Tensor, total: int, batch_dim: int = 0, group=None) -> Optional[torch.Tensor]:
    """Gather per-rank result slices (batch on `batch_dim`) into one tensor on rank 0, in image order.

    Shards may be uneven (see shard_range): every rank pads its slice to the largest shard so that a single
    fixed-size gather is issued, and rank 0 trims the padding.  Returns None on the other ranks."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [shard_range(total, world, r)[1] for r in range(world)]
    cap = max(counts) if counts else 0
    x = local.movedim(batch_dim, 0).contiguous()
    if x.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {x.shape[0]} images, expected {counts[rank]}")
    if x.shape[0] < cap:
        pad = torch.zeros((cap - x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        x = torch.cat([x, pad], dim=0)
    if world == 1:
        out = x[: counts[0]]
        return out.movedim(0, batch_dim)
    bufs: Optional[List[torch.Tensor]] = None
    if rank == 0:
        bufs = [torch.empty_like(x) for _ in range(world)]
    dist.gather(x, bufs, dst=0, group=group)
    if rank != 0:
        return None
    out = torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
    return out.movedim(0, batch_dim)


def gather_results(local: Dict[str, torch.Tensor], total: int, batch_dims: Dict[str, int], group=None
                   ) -> Optional[Dict[str, torch.Tensor]]:
    """Gather a dict of per-rank results (logits [b,C], avg_maps [L,b,N,N], cls_maps [L,b,H,N], rollout [b,n]...)."""
    out = {}
    for k in sorted(local):
        g = gather_to_rank0(local[k], total, batch_dims.get(k, 0), group)
        if g is not None:
            out[k] = g
    return out if dist.get_rank(group) == 0 else None


RESULT_BATCH_DIMS = {"logits": 0, "rollout": 0, "avg_maps": 1, "cls_maps": 1, "heads": 1, "hidden": 1}
