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


class RecordGatherer:
    """Pipelined gather of equal-sized record shards to `dst` (the steady state of a frame stream: SURVEY 8e,
    "issued per chunk on a side stream to overlap with compute").

    `depth` record buffers rotate: a step writes its records into `next_buffer()`, `gather()` starts the
    collective asynchronously (NCCL runs it on its own stream), and the buffer is only waited for when it comes
    round again - so the gather of step k overlaps the kernels of step k+1.  `flush()` waits for everything.
    """

    def __init__(self, n_local: int, record_bytes: int, device, dst: int = 0, group=None, depth: int = 2):
        self.group, self.dst, self.depth = group, dst, depth
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.send = [torch.empty((n_local, record_bytes), dtype=torch.uint8, device=device) for _ in range(depth)]
        self.out = [torch.empty((self.world, n_local, record_bytes), dtype=torch.uint8, device=device)
                    if self.rank == dst else None for _ in range(depth)]
        self.work = [None] * depth
        self.i = 0

    def next_buffer(self) -> torch.Tensor:
        slot = self.i % self.depth
        if self.work[slot] is not None:
            self.work[slot].wait()          # device-side dependency on NCCL, host-side wait on gloo
            self.work[slot] = None
        return self.send[slot]

    def gather(self):
        """Start gathering the buffer handed out by the last next_buffer(); returns on `dst` the
        [world * n_local, record_bytes] view that holds the result once the slot has been waited for."""
        slot = self.i % self.depth
        outs = list(self.out[slot].unbind(0)) if self.rank == self.dst else None
        self.work[slot] = dist.gather(self.send[slot], outs, dst=self.dst, group=self.group, async_op=True)
        self.i += 1
        return self.out[slot].view(-1, self.send[slot].shape[1]) if self.rank == self.dst else None

    def flush(self) -> None:
        for slot in range(self.depth):
            if self.work[slot] is not None:
                self.work[slot].wait()
                self.work[slot] = None
