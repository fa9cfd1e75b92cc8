#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/ (run in the build container).

  python scripts/profile_summary.py <round-tag> <launches.csv> <full.ncu-rep>
"""
import collections, csv, json, os, subprocess, sys

tag, launches_csv, rep = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: per-kernel totals and shares -------------------------------------------------
rows = [r for r in csv.reader(open(launches_csv)) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    unit = r[ui]
    v_us = v / 1000.0 if unit in ("ns", "nsecond") else v * 1000.0 if unit in ("ms", "msecond") else v
    name = r[ki].split("(")[0]
    tot[name] += v_us; cnt[name] += 1
total = sum(tot.values())
with open(os.path.join(out_dir, f"{tag}_launches_summary.md"), "w") as f:
    f.write(f"# {tag}: ncu launch list of `python bench.py --steps 20 --warmup 3 --no-cpu-baseline` (first 400 launches)\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` - per-launch times are cold-cache and serialised: compare SHARES.\n\n")
    f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write(f"| `{k[:90]}` | {cnt[k]} | {v:.1f} | {100 * v / total:.1f}% | {v / cnt[k]:.1f} |\n")
os.system(f"cp {launches_csv} {out_dir}/{tag}_launches.csv")

# ---- full capture of the top kernel --------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u, v = rr[0], rr[1], rr[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_elapsed", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
vals = {}
for i, name in enumerate(h):
    if name in want:
        vals[name] = (v[i], u[i])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
sr = list(csv.reader(src.splitlines()))
sh = sr[1]; idx = {x: i for i, x in enumerate(sh)}
stalls = [x for x in sh if x.startswith("stall_") and "Not Issued" not in x]
st = collections.Counter(); samples = 0
for r in sr[2:]:
    if len(r) < len(sh): continue
    samples += int(r[idx["# Samples"]] or 0)
    for s_ in stalls: st[s_] += int(r[idx[s_]] or 0)
def num(x):
    return float(x.replace(",", ""))
rd, wr = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum")
def to_bytes(val, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return num(val) * m.get(unit, 1)
traffic = to_bytes(*rd) + to_bytes(*wr)
with open(os.path.join(out_dir, f"{tag}_fused_tc_kernel.md"), "w") as f:
    f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` of the dominant kernel (cfg1, 256 frames per launch)\n\n")
    f.write("| metric | value | unit |\n|---|---:|---|\n")
    for k in want:
        if k in vals:
            f.write(f"| {k} | {vals[k][0][:110]} | {vals[k][1]} |\n")
    f.write(f"| DRAM traffic (read + write) per launch | {traffic / 1e6:.1f} | MB |\n")
    f.write(f"| algorithmic bytes per launch (256 x 6,554,752) | {256 * 6554752 / 1e6:.1f} | MB |\n")
    f.write(f"\nWarp-state samples: {samples}\n\n| stall reason | samples | share |\n|---|---:|---:|\n")
    for k, c in st.most_common(10):
        f.write(f"| {k} | {c} | {100 * c / max(samples, 1):.1f}% |\n")
tj = os.path.join(out_dir, "traffic.json")
d = json.load(open(tj)) if os.path.isfile(tj) else {}
d["cfg1"] = traffic
d["cfg1_source"] = f"{tag}_fused_tc_kernel.md (dram__bytes_read.sum + dram__bytes_write.sum, one launch of 256 frames)"
json.dump(d, open(tj, "w"), indent=1)
print("traffic", traffic / 1e6, "MB;", "wrote profiles for", tag)
