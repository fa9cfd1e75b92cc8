"""Frame sharding across GPUs: one process per GPU (torchrun), contiguous frame ranges, NO collective
inside the path; only the final per-frame records are gathered to rank 0 (NCCL over NVLink on the
GPU box, gloo in the CPU tests).  Binary masks are never gathered (north star)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of rank `rank`; the first n_frames % world ranks get one extra frame."""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_records(records: torch.Tensor, n_frames: int, dst: int = 0, group=None):
    """records: this rank's [n_local, record_bytes] u8 tensor -> on `dst` the [n_frames, record_bytes]
    tensor in frame order, elsewhere None.  Shards may be ragged (padded to the largest shard)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return records
    sizes = [shard_range(n_frames, r, world) for r in range(world)]
    max_local = max(hi - lo for lo, hi in sizes)
    rb = records.shape[1]
    send = records
    if records.shape[0] < max_local:
        send = torch.zeros((max_local, rb), dtype=records.dtype, device=records.device)
        send[:records.shape[0]] = records
    send = send.contiguous()
    if rank == dst:
        bufs = [torch.empty_like(send) for _ in range(world)]
        dist.gather(send, bufs, dst=dst, group=group)
        return torch.cat([bufs[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)], 0)
    dist.gather(send, None, dst=dst, group=group)
    return None
