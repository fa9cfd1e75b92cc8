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

    def wait_latest(self) -> None:
        """Make the view returned by the last gather() safe to read (the slot is reused `depth` steps later)."""
        slot = (self.i - 1) % self.depth
        if self.i and self.work[slot] is not None:
            self.work[slot].wait()
            self.work[slot] = None

    def flush(self) -> None:
        for slot in range(self.depth):
            if self.work[slot] is not None:
                self.work[slot].wait()
                self.work[slot] = None


def device_view(ptr: int, nbytes: int, device) -> torch.Tensor:
    """Zero-copy u8 tensor over raw device memory (used for buffers allocated through the C ABI)."""
    class _Raw:
        __cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(_Raw(), device=device)


class PeerRecordSink:
    """Records of every rank land in ONE buffer on rank `dst` without a collective: the buffer is allocated through the
    C ABI on `dst` (va_peer_alloc), its IPC handle is broadcast once, every other rank maps it (va_peer_open).  A step's
    records are written by the tail kernel into a LOCAL staging buffer (two rotate); a copy engine then moves them over
    NVLink into the rank's slot on `dst` (va_peer_put on a side stream, overlapped with the next step) and the rank's
    flag is raised behind the copy (va_signal) - no SM is spent on the gather and no NCCL kernel competes with the
    persistent mask kernel.  `dst` writes its own slot directly and can wait for a step with `wait(step)`.

    Layout on `dst`: [depth][world][n_local][record_bytes] then world int32 flags.  `depth` slots rotate (a frame
    stream keeps depth = number of chunks, i.e. every record has its own place)."""

    def __init__(self, eng, n_local: int, depth: int, group=None, dst: int = 0):
        self.eng, self.group, self.dst, self.depth, self.n_local = eng, group, dst, depth, n_local
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.rb = eng.record_bytes
        self.slot_bytes = self.world * n_local * self.rb
        self.flags_off = (depth * self.slot_bytes + 255) // 256 * 256
        total = self.flags_off + 4 * self.world
        box = [None]
        self.base, err = 0, ""
        if self.rank == dst:
            try:
                self.base, handle = eng.peer_alloc(total)
                box[0] = handle
            except Exception as e:  # noqa: BLE001
                err = str(e)
        dist.broadcast_object_list(box, src=dst, group=group)
        if self.rank != dst and box[0] is not None:
            try:
                self.base = eng.peer_open(box[0])
            except Exception as e:  # noqa: BLE001
                err = str(e)
        # every rank must agree on whether the mapping exists
        ok = torch.tensor([1 if self.base else 0], dtype=torch.int32, device=torch.device("cuda", eng.device))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            if self.base:
                (eng.peer_free if self.rank == dst else eng.peer_close)(self.base)
                self.base = 0
            raise RuntimeError("peer mapping of the record buffer failed on at least one rank: " + (err or "see other ranks"))
        self.total = total
        self.step = 0
        dev = torch.device("cuda", eng.device)
        self.remote = self.rank != dst
        if self.remote:
            self.stage = [torch.empty(n_local * self.rb, dtype=torch.uint8, device=dev) for _ in range(2)]
            self.copied = [None, None]
            self.side = torch.cuda.Stream(device=dev)

    def _slot_ptr(self) -> int:
        return self.base + (self.step % self.depth) * self.slot_bytes + self.rank * self.n_local * self.rb

    def records_ptr(self) -> int:
        """Address the tail kernel writes the records of the current step to: the rank's slot itself on `dst`, a local
        staging buffer elsewhere (waits, on the current stream, until the copy that last used it has finished)."""
        if not self.remote:
            return self._slot_ptr()
        s = self.step & 1
        if self.copied[s] is not None:
            torch.cuda.current_stream().wait_event(self.copied[s])
        return self.stage[s].data_ptr()

    def commit(self, n_frames: int | None = None) -> None:
        """Call after va_run_fused of the current step was enqueued: moves the step's records to `dst` (copy engine,
        side stream) and raises this rank's flag to step + 1 behind the copy."""
        flag = self.base + self.flags_off + 4 * self.rank
        if not self.remote:
            self.step += 1
            self.eng.signal(flag, self.step)
            return
        s = self.step & 1
        nbytes = (self.n_local if n_frames is None else n_frames) * self.rb
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream())
        self.side.wait_event(done)
        dst_ptr = self._slot_ptr()
        self.step += 1
        with torch.cuda.stream(self.side):
            self.eng.peer_put(dst_ptr, self.stage[s].data_ptr(), nbytes)
            self.eng.signal(flag, self.step)
            self.copied[s] = torch.cuda.Event()
            self.copied[s].record(self.side)

    def drain(self) -> None:
        """The current stream waits for this rank's outstanding copies (no-op on `dst`)."""
        if self.remote:
            torch.cuda.current_stream().wait_stream(self.side)

    def wait(self, step: int | None = None) -> None:
        """`dst` only: the current stream waits until every rank has committed `step` (default: the latest) steps."""
        self.eng.wait_flags(self.base + self.flags_off, self.world, self.step if step is None else step)

    def view(self, slot: int) -> torch.Tensor:
        """`dst` only: [world * n_local, record_bytes] u8 view of a slot."""
        v = device_view(self.base + slot * self.slot_bytes, self.slot_bytes, torch.device("cuda", self.eng.device))
        return v.view(self.world * self.n_local, self.rb)

    def close(self) -> None:
        if self.remote:
            self.side.synchronize()
        if self.rank == self.dst:
            self.eng.peer_free(self.base)
        else:
            self.eng.peer_close(self.base)
