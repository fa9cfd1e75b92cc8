"""CPU, world_size 2 over gloo: frame sharding + record gather (the N>1 host logic of bench.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vision_assist_b200.sharding import RecordGatherer, gather_records, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 256, 65536, 65537):
        for w in (1, 2, 3, 4, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in parts) - min(h - l for l, h in parts) <= 1


def _worker(rank, world, port, n_frames, rb, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_frames, rank, world)
    # each "record" encodes its global frame index so that order and raggedness are checked
    rec = torch.zeros((hi - lo, rb), dtype=torch.uint8)
    for k in range(hi - lo):
        rec[k] = torch.tensor([(lo + k + j) % 251 for j in range(rb)], dtype=torch.uint8)
    out = gather_records(rec, n_frames, dst=0)
    if rank == 0:
        want = np.array([[(f + j) % 251 for j in range(rb)] for f in range(n_frames)], np.uint8)
        q.put(bool(np.array_equal(out.numpy(), want)))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_records_world2_ragged():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, 48, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok


def _pipelined_worker(rank, world, port, n_local, rb, steps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = RecordGatherer(n_local, rb, "cpu", dst=0, depth=2)
    ok, views = True, []
    for step in range(steps):
        buf = g.next_buffer()
        buf.copy_(torch.full((n_local, rb), (7 * step + rank) % 256, dtype=torch.uint8))
        views.append((step, g.gather()))
        if rank == 0 and step >= 1:          # the previous step's slot is complete once it is waited for
            pass
    g.flush()
    if rank == 0:
        # the last `depth` steps are still resident in their slots
        for step, v in views[-2:]:
            want = np.concatenate([np.full((n_local, rb), (7 * step + r) % 256, np.uint8) for r in range(world)])
            ok = ok and bool(np.array_equal(v.numpy(), want))
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_pipelined_record_gatherer_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pipelined_worker, args=(r, 2, port, 5, 32, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok
