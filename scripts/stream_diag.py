"""Per-rank breakdown of the cfg3 stream with the peer record sink (developer diagnostic; torchrun, >= 2 GPUs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from vision_assist_b200 import synth
from vision_assist_b200.engine import MaskGridEngine
from vision_assist_b200.sharding import PeerRecordSink

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H = W = 640; n = 8; B = 256; chunks = int(os.environ.get("VA_CHUNKS", "32"))
eng = MaskGridEngine(H=H, W=W, mh=160, mw=160, max_n=n, gs=20, max_batch=B, device=local)
masks = torch.empty((B, n, H, W), dtype=torch.uint8, device="cuda")
local_rec = torch.empty((B, eng.record_bytes), dtype=torch.uint8, device="cuda")
gen = os.environ.get("VA_GEN", "0") == "1"
p, c, b, cnt = synth.make_batch_device(rank * 1000, B, n, H, W, 160, 160)
for mode in ("local", "sink_run_only", "sink_full", "local"):
    sink = PeerRecordSink(eng, B, depth=chunks) if mode != "local" else None
    run_ms = com_ms = 0.0
    evs = []
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for k in range(chunks):
        if gen:
            p, c, b, cnt = synth.make_batch_device(rank * 100000 + k * B, B, n, H, W, 160, 160)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        if sink is None:
            eng.run(p, c, b, cnt, masks_out=masks, records_out=local_rec)
        else:
            eng.run(p, c, b, cnt, masks_out=masks, records_ptr=sink.records_ptr())
        e1.record()
        if sink is not None and mode == "sink_full":
            sink.commit()
        e2.record()
        evs.append((e0, e1, e2))
    if sink is not None and mode == "sink_full":
        sink.drain()
        if rank == 0:
            sink.wait()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    run_ms = sum(a.elapsed_time(b_) for a, b_, _ in evs); com_ms = sum(b_.elapsed_time(c_) for _, b_, c_ in evs)
    print(f"[rank {rank}] {mode}: run {run_ms / chunks * 1e3:.1f} us/chunk, commit {com_ms / chunks * 1e3:.1f} us/chunk, wall {wall / chunks * 1e3:.1f} us/chunk", flush=True)
    dist.barrier()
    if sink is not None:
        sink.close()
dist.destroy_process_group()
