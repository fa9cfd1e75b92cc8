#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: stall reasons overall and the hottest SASS instructions."""
import csv, sys, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
tot = collections.Counter(); total_samples = 0; total_inst = 0
for r in data:
    if len(r) < len(hdr): continue
    total_samples += int(r[idx["# Samples"]] or 0); total_inst += int(r[idx["Instructions Executed"]] or 0)
    for s in stalls: tot[s] += int(r[idx[s]] or 0)
print(f"total samples {total_samples}  warp-instructions {total_inst}")
for s, v in tot.most_common(10): print(f"  {s:28s} {v:8d} {100*v/max(1,total_samples):5.1f}%")
print("--- hottest instructions (samples, execs, top stalls) ---")
order = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]] or 0) if len(data[i]) >= len(hdr) else 0)
for i in order[:topn]:
    r = data[i]
    st = sorted(((int(r[idx[s]] or 0), s) for s in stalls), reverse=True)[:3]
    print(f"{i:5d} {int(r[idx['# Samples']]):6d} {int(r[idx['Instructions Executed']]):9d}  {r[idx['Source']].strip()[:70]:70s} " + " ".join(f"{s[6:]}={v}" for v, s in st if v))
