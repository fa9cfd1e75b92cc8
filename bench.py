#!/usr/bin/env python
"""bench.py - frames/sec through mask -> grid -> penalty -> protrusion (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg1|cfg2|cfg3]

One "step" = one pass of the hot path over one batch of synthetic frames (cfg1: 256 frames of
640x640, prototypes 32x160x160, 8 instances, 20-px cells - BASELINE.json configs[1]).
  value        : whole-job frames/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e          : same metric through the host-buffer C-ABI call (pinned host tensors in, records back on
                 the host; H2D / D2H copies inside the timed region), next to the measured PCIe ceiling
  roofline     : algorithmic bytes of the dominant kernel / its CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline : the oracle port (reference algorithm) on this box's host cores, bounded sample of the same frames
  cfg2 / cfg3  : extra blocks of the default line - BASELINE configs[2] (1080p, 32 instances; N = 1 only) and
                 configs[3] (a 65,536-frame stream generated on the device, sharded over the ranks, frames/s with
                 and without landing every record on rank 0)
N > 1 (torchrun): every rank runs the same per-GPU batch on its own frames (weak scaling), no collective inside
the path; every step's records are moved by a copy engine into rank 0's peer-mapped buffer over NVLink (C ABI
va_peer_*, one flag per rank and step) inside the timed region; NCCL only for the rendezvous and the timing reduction.
--impl reference times the reference algorithm's CPU implementation (oracle port) instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec (mask->grid->penalty->protrusion)"
WORKLOADS = {
    # name: H, W, mh, mw, n, gs, B
    "cfg1": dict(H=640, W=640, mh=160, mw=160, n=8, gs=20, B=256,
                 desc="BASELINE configs[1]: batch of 256 synthetic 640x640 frames, protos 32x160x160, 8 instances, gs=20"),
    "cfg2": dict(H=1080, W=1920, mh=160, mw=160, n=32, gs=20, B=32,
                 desc="BASELINE configs[2]: 1920x1080 frames, 32 instances/frame, protos 32x160x160, gs=20"),
    "cfg3": dict(H=640, W=640, mh=160, mw=160, n=8, gs=20, B=256, stream=65536,
                 desc="BASELINE configs[3]: frame-sharded stream of 65,536 640x640 frames, chunks of 256, records landed on rank 0"),
}


def make_config(wl: dict) -> dict:
    """The `config` object of the JSON line - identical for --impl ours and --impl reference."""
    return {"workload": wl["desc"], "frames_per_step_per_gpu": wl["B"], "write_masks": True,
            "frame": f'{wl["W"]}x{wl["H"]}', "protos": f'32x{wl["mh"]}x{wl["mw"]}', "instances": wl["n"], "cell_px": wl["gs"],
            "inputs": "per-frame seeded synthetic frames (vision_assist_b200.synth, seed 0xB2000000 + frame index), "
                      "256 distinct frames per GPU",
            "l2": "inputs + masks of one step (1.68 GB at cfg1) exceed the 126 MB L2"}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """In-process NVML sampling (every ~2 ms) of SM clock and clock-event reasons during the timed region."""

    def __init__(self, gpu_index: int):
        self.rows, self.gpu, self.stop_flag, self.thread, self.err = [], gpu_index, False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, str(e)[:100]

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # older binding name
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.time(), sm, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self, t_begin: float = 0.0, t_end: float = 1e30) -> dict:
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {self.err}"], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        nv = self.nv
        inside = [(sm, rs) for (t, sm, rs) in self.rows if t_begin <= t <= t_end]
        rows = inside if inside else [(sm, rs) for (_, sm, rs) in self.rows]
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(k for k, b in bits.items() if any(rs & b for _, rs in rows))
        sm = [s for s, _ in rows]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(sm), "samples_in_timed_region": len(inside), "how": "in-process NVML, 2 ms period"}


def pin_to_gpu_numa_node(gpu_index: int) -> str:
    """Bind this process (and the pinned buffers it allocates from now on, first touch) to the CPUs NVML reports as
    closest to the GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(gpu_index))
        return f"{len(os.sched_getaffinity(0))} CPUs (NVML ideal affinity of GPU {gpu_index})"
    except Exception as e:  # noqa: BLE001
        return f"not pinned ({str(e)[:60]})"


# ---------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference path on the host cores
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One worker = one host core: synthesise its frames (untimed), run the reference route on them `reps` times,
    return the seconds spent in the route."""
    idxs, wl, reps, warm = args
    import torch
    torch.set_num_threads(1)
    from oracle import pipeline as opl
    from vision_assist_b200 import synth
    frames = [synth.make_frame(i, wl["n"], wl["H"], wl["W"], wl["mh"], wl["mw"]) for i in idxs]
    if warm:
        opl.frame_from_tensors(*frames[0], (wl["H"], wl["W"]), wl["gs"], "contour")
        return 0.0
    t0 = time.perf_counter()
    for _ in range(reps):
        for p, c, b in frames:
            opl.frame_from_tensors(p, c, b, (wl["H"], wl["W"]), wl["gs"], "contour")
    return time.perf_counter() - t0


def cpu_baseline(wl: dict, frames_per_core: int, cores: int | None = None, steps: int = 1, warmup: int = 0) -> dict:
    """process_mask -> masks2segments -> scale_coords -> grid -> penalties -> peaks (reference route, real OpenCV
    contours) on `cores` processes, one torch thread each.  The frames are the GPU arm's (indices 0 .. B-1 of rank 0),
    dealt round-robin to the cores and repeated until every core has `frames_per_core` frames per step; a step is
    timed as the slowest worker's time in the route (inputs are synthesised before the clock starts, as on the GPU
    side); value = frames of the `steps` timed steps / the sum of their times."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    B = wl["B"]
    per_core = max(1, min(frames_per_core, (B + cores - 1) // cores))
    reps = max(1, frames_per_core // per_core)
    shards = [[(r + k * cores) % B for k in range(per_core)] for r in range(cores)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(shards[r][:1], wl, 1, True) for r in range(cores)])       # spawn + import warm-up
        for _ in range(warmup):
            pool.map(_cpu_worker, [(shards[r], wl, reps, False) for r in range(cores)])
        secs = 0.0
        for _ in range(steps):
            secs += max(pool.map(_cpu_worker, [(shards[r], wl, reps, False) for r in range(cores)]))
    total = per_core * reps * cores * steps
    return {"value": total / secs, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{steps} step(s) of {per_core * reps * cores} frames ({per_core} distinct frames per core x {reps}, the GPU "
                      f"arm's frame indices 0..{B - 1}), oracle port of the reference route (process_mask, findContours, "
                      f"contourArea, fillPoly grid, penalties, peaks), {secs:.2f} s"}


def run_reference(args, wl):
    """Reference arm: the reference algorithm's CPU implementation (oracle port) on every host core.  Exactly
    --steps timed steps after --warmup untimed ones; the per-step sample is sized so that the run stays within
    about two minutes (at most 48 frames per core per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    budget_s, sec_per_frame = 90.0, 0.025
    fpc = int(budget_s / ((steps + warmup) * sec_per_frame))
    fpc = max(1, min(48, fpc))
    base = cpu_baseline(wl, fpc, cores, steps=steps, warmup=warmup)
    v = base["value"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": 1000.0 * fpc * cores / v,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": make_config(wl),
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def _events():
    import torch
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def measure_device_resident(eng, tensors, masks, records, steps, warmup, sink=None, sampler=None, gatherer=None):
    """`steps` timed steps of va_run_fused on device-resident inputs.  With a PeerRecordSink every step's records go
    to rank 0 over NVLink and rank 0 waits for all flags inside the timed region.  -> dict(ms, kernel_ms, tail_ms, calls)"""
    import torch
    import torch.distributed as dist
    protos, coefs, boxes, counts = tensors
    # diagnostic only (VA_BENCH_COMMIT_EVERY=k): move the records of every k-th step only - what the per-step put costs
    commit_every = max(1, int(os.environ.get("VA_BENCH_COMMIT_EVERY", "1")))
    state = {"i": 0}

    def step():
        if gatherer is not None:          # fallback: NCCL gather of every step's records, overlapped with the next step
            buf = gatherer.next_buffer()
            eng.run(protos, coefs, boxes, counts, masks_out=masks, records_out=buf, write_masks=masks is not None)
            gatherer.gather()
        elif sink is None:
            eng.run(protos, coefs, boxes, counts, masks_out=masks, records_out=records, write_masks=masks is not None)
        else:
            eng.run(protos, coefs, boxes, counts, masks_out=masks, write_masks=masks is not None, records_ptr=sink.records_ptr())
            state["i"] += 1
            if state["i"] % commit_every == 0:
                sink.commit()

    for _ in range(max(warmup, 3)):
        step()
    if sink is not None:
        sink.drain()
    if sink is not None and sink.rank == sink.dst:
        sink.wait()
    if gatherer is not None:
        gatherer.flush()
    torch.cuda.synchronize()
    if sink is not None or gatherer is not None:
        dist.barrier()
    eng.profile(8 if steps >= 64 else 1)
    e0, e1 = _events()
    torch.cuda.synchronize()
    t_begin = time.time()
    e0.record()
    for _ in range(steps):
        step()
    if sink is not None:
        sink.drain()                       # this rank's copies to rank 0 are inside its clock
        if sink.rank == sink.dst:
            sink.wait()                    # every rank's records of the last step have landed on rank 0
    if gatherer is not None:
        gatherer.flush()
    e1.record()
    torch.cuda.synchronize()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    asm_ms, tail_ms, calls = eng.profile_read()
    eng.profile(False)
    if sink is not None or gatherer is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return dict(ms=ms, kernel_ms=asm_ms / max(calls, 1), tail_ms=tail_ms / max(calls, 1), calls=calls,
                t_begin=t_begin, t_end=t_end, launches_per_step=eng.last_launch_count + (1 if sink is not None else 0))


def pcie_ceiling(nbytes: int, world: int, reps: int = 4) -> float:
    """Plain pinned cudaMemcpyAsync H2D bandwidth (GB/s per rank) with all ranks copying at the same time."""
    import torch
    import torch.distributed as dist
    src = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s = float(t.item())
    return nbytes * reps / s / 1e9


def run_stream(eng, wl, rank, world, sink_factory):
    """BASELINE configs[3]: 65,536 frames, contiguous shards of 65,536 / world, chunks of 256 generated on the device
    from the per-frame seeds right before they are processed (generation is outside the timing events: the timed
    region is the hot path only, and no byte crosses PCIe).  Two passes: records stay on the rank / records land on
    rank 0 (peer stores, final wait on rank 0 inside the clock).  -> dict"""
    import torch
    import torch.distributed as dist
    from vision_assist_b200 import synth
    from vision_assist_b200.sharding import shard_range
    H, W, mh, mw, n, B, total = (wl[k] for k in ("H", "W", "mh", "mw", "n", "B", "stream"))
    lo, hi = shard_range(total, rank, world)
    n_chunks = (hi - lo + B - 1) // B
    masks = torch.empty((B, n, H, W), dtype=torch.uint8, device="cuda")
    local = torch.empty((B, eng.record_bytes), dtype=torch.uint8, device="cuda")
    out = {"frames": total, "chunk": B, "frames_per_rank": hi - lo,
           "generator": "on-device, hash of (0xB2000000 + frame index, draw): independent of chunking and sharding"}
    non_simple = 0
    for mode in ("no_gather", "gather"):
        sink = sink_factory(n_chunks) if (mode == "gather" and world > 1) else None
        nccl_fallback = mode == "gather" and world > 1 and sink is None
        allrec = (torch.empty((n_chunks * B, eng.record_bytes), dtype=torch.uint8, device="cuda")
                  if (mode == "gather" and (world == 1 or nccl_fallback)) else None)
        ms = 0.0
        evs = []
        for k in range(n_chunks):
            f0 = lo + k * B
            nb = min(B, hi - f0)
            p, c, b, cnt = synth.make_batch_device(f0, nb, n, H, W, mh, mw)
            e0, e1 = _events()
            e0.record()
            if sink is not None:
                eng.run(p, c, b, cnt, masks_out=masks[:nb], write_masks=True, records_ptr=sink.records_ptr())
                sink.commit(nb)
            elif allrec is not None:
                eng.run(p, c, b, cnt, masks_out=masks[:nb], records_out=allrec[k * B:k * B + nb], write_masks=True)
            else:
                eng.run(p, c, b, cnt, masks_out=masks[:nb], records_out=local[:nb], write_masks=True)
            e1.record()
            evs.append((e0, e1))
            if mode == "no_gather" and k % 32 == 0:
                non_simple += int((local[:nb, 0] & 8).ne(0).sum().item())
        if sink is not None:
            e0, e1 = _events()
            e0.record()
            sink.drain()                   # this rank's copies to rank 0 have completed: inside its clock
            e1.record()
            evs.append((e0, e1))
            # the chunks are generated (untimed, ~10 ms each) between the timed regions, so the ranks drift apart by
            # milliseconds: line them up before rank 0 checks the flags, or that drift would be billed to the gather
            torch.cuda.synchronize()
            dist.barrier()
            if rank == 0:
                e0, e1 = _events()
                e0.record()
                sink.wait()
                e1.record()
                evs.append((e0, e1))
        if nccl_fallback:                  # one NCCL gather of the whole shard at the end of the stream
            from vision_assist_b200.sharding import gather_records
            e0, e1 = _events()
            e0.record()
            gather_records(allrec[:hi - lo], total, dst=0)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b_) for a, b_ in evs)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        out[f"frames_per_s_{mode}"] = total / (ms / 1000.0)
        out[f"ms_{mode}"] = ms
        if sink is not None:
            if rank == 0:
                flags = sink.view(0)[:, 0]            # header.flags byte 0 of every record of chunk 0
                out["records_on_rank0_chunk0"] = int(flags.numel())
            dist.barrier()
            sink.close()
    out["non_simple_frames_sampled"] = non_simple
    out["gather"] = ("every chunk's records are put into rank 0's peer-mapped buffer by a copy engine over NVLink (side stream), one "
                     "flag per rank and chunk; rank 0 waits for the last flags inside the clock") if world > 1 else \
        "single GPU: records written to the full-stream buffer"
    return out


def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    from vision_assist_b200 import synth
    from vision_assist_b200.engine import MaskGridEngine
    from vision_assist_b200.sharding import PeerRecordSink, RecordGatherer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    affinity = pin_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    H, W, mh, mw, n, gs, B = (wl[k] for k in ("H", "W", "mh", "mw", "n", "gs", "B"))
    eng = MaskGridEngine(H=H, W=W, mh=mh, mw=mw, max_n=n, gs=gs, max_batch=B, device=local,
                         tensor_core=not args.no_tensor_core)

    sink_note = {}

    def sink_factory(depth):
        """Peer-mapped record buffer on rank 0; None (-> NCCL gather, round-1 path) when the mapping is not available."""
        try:
            return PeerRecordSink(eng, B, depth=depth)
        except Exception as e:  # noqa: BLE001
            sink_note["fallback"] = f"peer mapping unavailable ({str(e)[:80]}): records gathered with NCCL instead"
            return None

    if args.workload == "cfg3":
        res = run_stream(eng, wl, rank, world, sink_factory)
        if rank == 0:
            line = {"metric": METRIC, "value": res["frames_per_s_gather"], "unit": "frames/s", "n_gpus": world,
                    "steps": res["frames"] // (B * world), "warmup": 0, "ms_per_step": res["ms_gather"] / max(1, res["frames"] // (B * world)),
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": "f32 (tf32x3 tensor-core contraction, fp32 blend, f64 penalties, u8 masks)", "data": "synthetic",
                    "config": make_config(wl), "cfg3": res}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # B distinct per-frame-seeded frames per rank (SURVEY 8d generator), generated on the host once, resident in HBM
    hp, hc, hb, hn = synth.make_batch(rank * B, B, n, H, W, mh, mw, max_n=n)
    tensors = tuple(t.cuda() for t in (hp, hc, hb, hn))
    masks = torch.empty((B, n, H, W), dtype=torch.uint8, device="cuda")
    records = torch.empty((B, eng.record_bytes), dtype=torch.uint8, device="cuda")
    in_bytes = hp.numel() * 4 + hc.numel() * 4 + hb.numel() * 4

    sink = sink_factory(2) if world > 1 else None
    gatherer = RecordGatherer(B, eng.record_bytes, "cuda", dst=0, depth=2) if (world > 1 and sink is None) else None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    m = measure_device_resident(eng, tensors, masks, records, args.steps, args.warmup, sink, sampler, gatherer)
    clocks = sampler.stop(m["t_begin"], m["t_end"]) if rank == 0 else None
    ms = m["ms"]
    value = B * world * args.steps / (ms / 1000.0)
    if sink is not None:
        dist.barrier()
        sink.close()
    non_simple = int((records[:, 0] & 8).ne(0).sum().item()) if sink is None else None

    # grid-only mode (u8 masks never written; bit-packed copy for the contour step), reported separately
    g = measure_device_resident(eng, tensors, None, records, max(5, min(args.steps, 50)), 3)
    grid_only = B * max(5, min(args.steps, 50)) / (g["ms"] / 1000.0)

    # end to end through the host-buffer C-ABI call (pinned host memory in, records back on the host)
    pp, pc, pb, pn = (t.pin_memory() for t in (hp, hc, hb, hn))
    hrec = torch.empty((B, eng.record_bytes), dtype=torch.uint8, pin_memory=True)
    e2e_steps = max(2, min(args.steps, 5))
    eng.run_host(pp, pc, pb, pn, records_out=hrec)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.run_host(pp, pc, pb, pn, records_out=hrec)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = B * world * e2e_steps / e2e_s
    # separately labelled: the same call with fp16 prototypes in host memory (model run with half=True; the reference
    # computes on protos.float(), ops.py:724).  Different input values than the fp32 line - not the headline.
    e2e_f16 = None
    if not args.no_extras:
        ph = hp.half().pin_memory()
        eng.run_host(ph, pc, pb, pn, records_out=hrec)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.run_host(ph, pc, pb, pn, records_out=hrec)
        f16_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([f16_s], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            f16_s = float(t.item())
        e2e_f16 = {"value": B * world * e2e_steps / f16_s, "unit": "frames/s",
                   "h2d_bytes_per_step": hp.numel() * 2 + hc.numel() * 4 + hb.numel() * 4 + B * 4,
                   "note": "va_run_fused_host_f16: fp16 prototypes in pinned host memory, widened exactly on the device "
                           "(= protos.float()), then the fp32 path; inputs are the fp32 line's prototypes rounded to fp16"}
        del ph
    ceiling_gbs = pcie_ceiling(256 << 20, world)
    e2e_ceiling = world * ceiling_gbs * 1e9 / ((in_bytes + B * 4) / B)      # frames/s if the copies ran at the ceiling

    cfg3 = run_stream(eng, WORKLOADS["cfg3"], rank, world, sink_factory) if (args.workload == "cfg1" and not args.no_extras) else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    latency = measure_latency(wl, local) if world == 1 else None
    peak, peak_src = load_peaks()
    alg_kernel = B * (4 * 32 * mh * mw + 4 * n * 32 + 16 * n + n * H * W)       # dominant kernel: assembly
    alg_step = B * eng.algorithmic_bytes_per_frame(n, True)
    kernel_ms = m["kernel_ms"]
    achieved = alg_kernel / (kernel_ms / 1000.0) / 1e9
    ncu_traffic = None
    tfile = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tfile):
        try:
            ncu_traffic = json.load(open(tfile)).get(args.workload)
        except Exception:
            ncu_traffic = None
    cfg = make_config(wl)
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (tf32x3 tensor-core contraction, fp32 blend, f64 penalties, u8 masks)",
        "data": "synthetic", "config": cfg,
        "notes": {"contraction": "tcgen05" if eng.uses_tensor_core else "cuda-core",
                  "multi_gpu": "frames sharded per rank, no collective in the path; every step's records are moved by a copy engine "
                               "into rank 0's peer-mapped buffer over NVLink (side stream, one flag per rank and step), rank 0 waits "
                               "for the last flags inside the timed region" if world > 1 else "single GPU",
                  "cpu_affinity": affinity,
                  "non_simple_frames_per_step": non_simple,
                  **sink_note,
                  "contour_step": "exact: every record is built from the polygon the reference keeps (findContours RETR_EXTERNAL, "
                                  "most points, contourArea selection); non_simple = frames whose selected mask is not one hole-free blob"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": in_bytes + B * 4,
                "d2h_bytes_per_step": B * eng.record_bytes, "steps": e2e_steps,
                "pcie_ceiling": {"h2d_gbs_per_gpu_all_ranks_concurrent": ceiling_gbs, "frames_per_s_at_ceiling": e2e_ceiling,
                                 "frac_of_ceiling": e2e_value / e2e_ceiling,
                                 "how": "pinned cudaMemcpyAsync of 256 MB x4, all ranks at once, slowest rank"},
                "note": "va_run_fused_host: pinned host tensors in, records in host memory out; grid-only instantiation (h_masks_out = "
                        "NULL: the u8 masks are neither written nor returned); PCIe-bound: 3.3 MB of fp32 prototypes per frame"},
        "e2e_fp16_protos": e2e_f16,
        "gpu_launches": m["launches_per_step"] * args.steps,
        "roofline": {"bound": "hbm", "kernel": "fused_tc_kernel" if eng.uses_tensor_core else "logits+upsample",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic, "algorithmic_bytes_per_launch": alg_kernel, "kernel_ms": kernel_ms,
                     "timed_launches": m["calls"],
                     "tail_ms": m["tail_ms"], "peak_source": peak_src,
                     "step_frac": (alg_step / (ms / args.steps / 1000.0) / 1e9) / peak},
        "grid_only_frames_per_s": grid_only,
        "latency_b1": latency,
    }
    if cfg3 is not None:
        line["cfg3"] = cfg3
    if world == 1 and args.workload == "cfg1" and not args.no_extras:
        line["cfg2"] = measure_cfg2(local, peak)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(wl, frames_per_core=768)      # about 10-15 s of work on every host core
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_cfg2(device, peak, steps: int = 20):
    """BASELINE configs[2] as an extra block of the N = 1 line: 32 frames of 1920x1080, 32 instances per frame."""
    import torch
    from vision_assist_b200 import synth
    from vision_assist_b200.engine import MaskGridEngine
    wl = WORKLOADS["cfg2"]
    H, W, mh, mw, n, gs, B = (wl[k] for k in ("H", "W", "mh", "mw", "n", "gs", "B"))
    eng = MaskGridEngine(H=H, W=W, mh=mh, mw=mw, max_n=n, gs=gs, max_batch=B, device=device)
    hp, hc, hb, hn = synth.make_batch(500_000, B, n, H, W, mh, mw, max_n=n)
    tensors = tuple(t.cuda() for t in (hp, hc, hb, hn))
    masks = torch.empty((B, n, H, W), dtype=torch.uint8, device="cuda")
    records = torch.empty((B, eng.record_bytes), dtype=torch.uint8, device="cuda")
    sampler = ClockSampler(device)
    sampler.start()
    m = measure_device_resident(eng, tensors, masks, records, steps, 3)
    clocks = sampler.stop(m["t_begin"], m["t_end"])
    alg_kernel = B * (4 * 32 * mh * mw + 4 * n * 32 + 16 * n + n * H * W)
    alg_step = B * eng.algorithmic_bytes_per_frame(n, True)
    out = {"config": make_config(wl), "value": B * steps / (m["ms"] / 1000.0), "unit": "frames/s", "steps": steps,
           "ms_per_step": m["ms"] / steps, "contraction": "tcgen05" if eng.uses_tensor_core else "cuda-core",
           "roofline": {"bound": "hbm", "kernel": "mask assembly (all launches of the step)",
                        "achieved": alg_kernel / (m["kernel_ms"] / 1000.0) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg_kernel / (m["kernel_ms"] / 1000.0) / 1e9 / peak, "kernel_ms": m["kernel_ms"],
                        "tail_ms": m["tail_ms"], "algorithmic_bytes_per_launch": alg_kernel,
                        "step_frac": (alg_step / (m["ms"] / steps / 1000.0) / 1e9) / peak},
           "non_simple_frames_per_step": int((records[:, 0] & 8).ne(0).sum().item()), "clocks": clocks}
    eng.close()
    del masks, tensors
    torch.cuda.empty_cache()
    return out


def measure_latency(wl, device, iters: int = 1000):
    """p50 / p99 single-frame latency (B = 1): launch -> record visible on the host (pinned D2H included),
    grid-only mode (the reference's FrameProcessor consumes only the grid), with and without a CUDA graph."""
    import torch
    from vision_assist_b200 import synth
    from vision_assist_b200.engine import MaskGridEngine
    H, W, mh, mw, n, gs = (wl[k] for k in ("H", "W", "mh", "mw", "n", "gs"))
    eng = MaskGridEngine(H=H, W=W, mh=mh, mw=mw, max_n=n, gs=gs, max_batch=1, device=device)
    hp, hc, hb, hn = synth.make_batch(424242, 1, n, H, W, mh, mw, max_n=n)
    protos, coefs, boxes, counts = hp.cuda(), hc.cuda(), hb.cuda(), hn.cuda()
    rec = torch.empty((1, eng.record_bytes), dtype=torch.uint8, device="cuda")
    hrec = torch.empty((1, eng.record_bytes), dtype=torch.uint8, pin_memory=True)
    out = {}
    eng.run(protos, coefs, boxes, counts, records_out=rec, write_masks=False)     # allocates the bit-mask scratch
    torch.cuda.synchronize()

    def once():
        eng.run(protos, coefs, boxes, counts, records_out=rec, write_masks=False)
        hrec.copy_(rec, non_blocking=True)
        torch.cuda.synchronize()

    def stats(fn):
        for _ in range(20):
            fn()
        ts = []
        for _ in range(iters):
            t0 = time.perf_counter()
            fn()
            ts.append((time.perf_counter() - t0) * 1e6)
        ts.sort()
        return {"p50_us": ts[len(ts) // 2], "p99_us": ts[int(len(ts) * 0.99) - 1]}

    out["eager"] = stats(once)
    try:
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for _ in range(3):
                eng.run(protos, coefs, boxes, counts, records_out=rec, write_masks=False)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            eng.run(protos, coefs, boxes, counts, records_out=rec, write_masks=False)
            hrec.copy_(rec, non_blocking=True)

        def replay():
            graph.replay()
            torch.cuda.synchronize()
        out["cuda_graph"] = stats(replay)
    except Exception as e:  # graph capture is an optimisation of the measurement, not of the path
        out["cuda_graph"] = {"error": str(e)[:120]}
    out["what"] = "B=1, grid-only, launch -> record in pinned host memory (perf_counter around run + D2H + sync)"
    # the drop-in FrameProcessor.__call__ (reference FrameProcessor.py:301-360) on the same frame: model shim returning the
    # head tensors already on the GPU, record -> peaks -> array A* on the host; Grid objects are built on first access only
    try:
        import numpy as np
        from vision_assist_b200.FrameProcessor import FrameProcessor, HeadOutputModel
        p1, c1, b1 = protos[0], coefs[0, :n].contiguous(), boxes[0, :n].contiguous()
        FrameProcessor._instance, FrameProcessor._initialized = None, False
        fp = FrameProcessor(HeadOutputModel(lambda frame: (p1, c1, b1)))
        frame = np.zeros((H, W, 3), np.uint8)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            lazy = stats(lambda: fp(frame))
            full = stats(lambda: (fp(frame), fp.grids))
        out["dropin_call"] = {"p50_us": lazy["p50_us"], "p99_us": lazy["p99_us"],
                              "with_grid_objects_p50_us": full["p50_us"],
                              "what": "FrameProcessor(frame) -> peaks + A* paths (array port), no Grid objects; "
                                      "with_grid_objects also reads .grids (the reference's pydantic object view)"}
        FrameProcessor._instance, FrameProcessor._initialized = None, False
    except Exception as e:
        out["dropin_call"] = {"error": str(e)[:160]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg1", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg2 / cfg3 blocks of the default line")
    ap.add_argument("--no-tensor-core", action="store_true", help="debug: CUDA-core contraction path")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
