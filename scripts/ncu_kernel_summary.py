#!/usr/bin/env python
"""One-kernel summary of an `ncu --set full` report (run in the build container):

  python scripts/ncu_kernel_summary.py <report.ncu-rep> <out.md> "<title>" [algorithmic MB per launch]
"""
import csv, subprocess, sys

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
alg = float(sys.argv[4]) if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u, v = rr[0], rr[1], rr[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
vals = {n: (v[i], u[i]) for i, n in enumerate(h) if n in want}


def to_bytes(val, unit):
    return float(val.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(val, unit):
    return float(val.replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(unit, 1)


traffic = to_bytes(*vals["dram__bytes_read.sum"]) + to_bytes(*vals["dram__bytes_write.sum"])
us = to_us(*vals["gpu__time_duration.sum"])
with open(out, "w") as f:
    f.write(f"# {title}\n\n`ncu --set full --clock-control none` (cold caches, serialised: compare with the CUDA-event times of bench.py / sweep)\n\n")
    f.write("| metric | value | unit |\n|---|---:|---|\n")
    for k in want:
        if k in vals:
            f.write(f"| {k} | {vals[k][0][:100]} | {vals[k][1]} |\n")
    f.write(f"| DRAM traffic (read + write) per launch | {traffic / 1e6:.1f} | MB |\n")
    f.write(f"| DRAM traffic / duration | {traffic / us / 1e3:.0f} | GB/s |\n")
    if alg:
        f.write(f"| algorithmic bytes per launch | {alg:.1f} | MB |\n| algorithmic bytes / duration | {alg * 1e6 / us / 1e3:.0f} | GB/s |\n")
print(out, f"{us:.1f} us, {traffic / 1e6:.1f} MB")
