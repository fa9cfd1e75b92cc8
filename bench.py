#!/usr/bin/env python
"""bench.py - frames/sec through mask -> grid -> penalty -> protrusion (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg1|cfg2]

One "step" = one pass of the hot path over one batch of synthetic frames (cfg1: 256 frames of
640x640, prototypes 32x160x160, 8 instances, 20-px cells - BASELINE.json configs[1]).
  value  : whole-job frames/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e    : same metric through the host-buffer C-ABI call (pinned host tensors in, records back on
           the host; H2D / D2H copies inside the timed region)
  roofline : algorithmic bytes of the dominant kernel / its CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline : the oracle port (reference algorithm) on this box's host cores, bounded sample
N > 1 (torchrun): every rank runs the same per-GPU batch on its own frames (weak scaling), no
collective inside the path, records gathered to rank 0 with NCCL inside the timed region.
--impl reference times the reference algorithm's CPU implementation (oracle port) instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
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
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t_begin: float = 0.0, t_end: float = 1e30) -> dict:
        """Summarise the samples taken inside [t_begin, t_end] (host clock); all samples if none fall inside."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        inside = [r for (t, r) in self.rows if t_begin <= t <= t_end + 0.05]
        rows = inside if inside else [r for (_, r) in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": len(inside)}


# ---------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference path on the host cores
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One worker = one host core: synthesise `count` frames (untimed), run the reference route on them, return
    the seconds spent in the route."""
    first, count, wl, warm = args
    import torch
    torch.set_num_threads(1)
    from oracle import pipeline as opl
    from vision_assist_b200 import synth
    frames = [synth.make_frame(first + i, wl["n"], wl["H"], wl["W"], wl["mh"], wl["mw"]) for i in range(count)]
    if warm:
        opl.frame_from_tensors(*frames[0], (wl["H"], wl["W"]), wl["gs"], "contour")
    t0 = time.perf_counter()
    for p, c, b in frames:
        opl.frame_from_tensors(p, c, b, (wl["H"], wl["W"]), wl["gs"], "contour")
    return time.perf_counter() - t0


def cpu_baseline(wl: dict, frames_per_core: int, cores: int | None = None, steps: int = 1, warmup: int = 0) -> dict:
    """process_mask -> masks2segments -> scale_coords -> grid -> penalties -> peaks (reference route) on `cores`
    processes, one torch thread each, disjoint frame shards.  A step = `frames_per_core` frames on every core, timed
    as the slowest worker's time in the route (inputs are synthesised before the clock starts, as on the GPU side);
    value = frames of the `steps` timed steps / the sum of their times."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(10_000 + 7 * r, 1, wl, True) for r in range(cores)])       # spawn + import warm-up
        for w in range(warmup):
            pool.map(_cpu_worker, [(15_000 + (w * cores + r) * frames_per_core, frames_per_core, wl, False) for r in range(cores)])
        secs = 0.0
        for k in range(steps):
            secs += max(pool.map(_cpu_worker, [(20_000 + (k * cores + r) * frames_per_core, frames_per_core, wl, False)
                                               for r in range(cores)]))
    total = frames_per_core * cores * steps
    return {"value": total / secs, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{steps} step(s) of {frames_per_core * cores} frames of the workload ({frames_per_core} per core), oracle "
                      f"port of the reference route (process_mask, findContours, fillPoly grid, penalties, peaks), {secs:.2f} s"}


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
            "data": "synthetic", "config": {"workload": wl["desc"], "frames_per_step": fpc * cores},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    from vision_assist_b200 import synth
    from vision_assist_b200.engine import MaskGridEngine
    from vision_assist_b200.sharding import RecordGatherer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    H, W, mh, mw, n, gs, B = (wl[k] for k in ("H", "W", "mh", "mw", "n", "gs", "B"))
    eng = MaskGridEngine(H=H, W=W, mh=mh, mw=mw, max_n=n, gs=gs, max_batch=B, device=local,
                         tensor_core=not args.no_tensor_core)

    # distinct synthetic frames per rank; generated on the host once, resident in HBM before timing
    uniq = min(B, 64)
    hp, hc, hb, hn = synth.make_batch(rank * 100_000, uniq, n, H, W, mh, mw, max_n=n)
    reps = (B + uniq - 1) // uniq
    protos = hp.repeat(reps, 1, 1, 1)[:B].contiguous().cuda()
    coefs = hc.repeat(reps, 1, 1)[:B].contiguous().cuda()
    boxes = hb.repeat(reps, 1, 1)[:B].contiguous().cuda()
    counts = hn.repeat(reps)[:B].contiguous().cuda()
    masks = torch.empty((B, n, H, W), dtype=torch.uint8, device="cuda")
    records = torch.empty((B, eng.record_bytes), dtype=torch.uint8, device="cuda")
    in_bytes = protos.numel() * 4 + coefs.numel() * 4 + boxes.numel() * 4

    # N > 1: the records of step k are gathered to rank 0 (NCCL) while step k+1 computes: two record buffers rotate
    gatherer = RecordGatherer(B, eng.record_bytes, "cuda", dst=0, depth=2) if world > 1 else None

    def step():
        if gatherer is None:
            eng.run(protos, coefs, boxes, counts, masks_out=masks, records_out=records, write_masks=True)
            return records
        buf = gatherer.next_buffer()
        eng.run(protos, coefs, boxes, counts, masks_out=masks, records_out=buf, write_masks=True)
        return gatherer.gather()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    eng.profile(8 if args.steps >= 64 else 1)   # kernel timing events on every 8th step of the timed region
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_begin = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    if gatherer is not None:
        gatherer.flush()                   # every gather finishes inside the timed region
    e1.record()
    torch.cuda.synchronize()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    launches = eng.last_launch_count * args.steps
    asm_ms, tail_ms, calls = eng.profile_read()
    eng.profile(False)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    value = B * world * args.steps / (ms / 1000.0)

    # grid-only mode (masks never written), reported separately
    for _ in range(3):
        eng.run(protos, coefs, boxes, counts, records_out=records, write_masks=False)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        eng.run(protos, coefs, boxes, counts, records_out=records, write_masks=False)
    g1.record()
    torch.cuda.synchronize()
    grid_only = B * args.steps / (g0.elapsed_time(g1) / 1000.0)

    # end to end through the host-buffer C-ABI call (pinned host memory in, records back on the host)
    pp, pc, pb, pn = (t.pin_memory() for t in (hp.repeat(reps, 1, 1, 1)[:B].contiguous(), hc.repeat(reps, 1, 1)[:B].contiguous(),
                                               hb.repeat(reps, 1, 1)[:B].contiguous(), hn.repeat(reps)[:B].contiguous()))
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    latency = measure_latency(wl, local) if world == 1 else None
    peak, peak_src = load_peaks()
    alg_kernel = B * (4 * 32 * mh * mw + 4 * n * 32 + 16 * n + n * H * W)       # dominant kernel: assembly
    alg_step = B * eng.algorithmic_bytes_per_frame(n, True)
    kernel_ms = asm_ms / max(calls, 1)
    achieved = alg_kernel / (kernel_ms / 1000.0) / 1e9
    ncu_traffic = None
    tfile = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tfile):
        try:
            ncu_traffic = json.load(open(tfile)).get(args.workload)
        except Exception:
            ncu_traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (tf32x3 tensor-core contraction, fp32 blend, f64 penalties, u8 masks)",
        "data": "synthetic",
        "config": {"workload": wl["desc"], "frames_per_step_per_gpu": B, "write_masks": True,
                   "l2": f"inputs {in_bytes / 1e6:.0f} MB + masks {masks.numel() / 1e6:.0f} MB per step exceed the 126 MB L2",
                   "contraction": "tcgen05" if eng.uses_tensor_core else "cuda-core",
                   "multi_gpu": "frames sharded per rank, no collective in the path; NCCL gather of every step's records to rank 0 inside the timed region, overlapped with the next step"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": in_bytes + B * 4,
                "d2h_bytes_per_step": B * eng.record_bytes, "steps": e2e_steps,
                "note": "va_run_fused_host: pinned host tensors in, records in host memory out (PCIe-bound)"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "fused_tc_kernel" if eng.uses_tensor_core else "logits+upsample",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic, "algorithmic_bytes_per_launch": alg_kernel, "kernel_ms": kernel_ms,
                     "timed_launches": calls,
                     "tail_ms": tail_ms / max(calls, 1), "peak_source": peak_src,
                     "step_frac": (alg_step / (ms / args.steps / 1000.0) / 1e9) / peak},
        "grid_only_frames_per_s": grid_only,
        "latency_b1": latency,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(wl, frames_per_core=96)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg1", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tensor-core", action="store_true", help="debug: CUDA-core contraction path")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
