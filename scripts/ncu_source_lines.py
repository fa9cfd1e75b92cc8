#!/usr/bin/env python
"""Warp-state samples of an `ncu --set full --import-source on` report attributed to CUDA source lines (build
container): the SASS of the report is matched, instruction by instruction, with `nvdisasm --print-line-info` of the
object file that was profiled.

  python scripts/ncu_source_lines.py <report.ncu-rep> <object.o> <mangled-kernel-substring> <out.md> "<title>" [top]
"""
import collections, csv, glob, os, re, subprocess, sys, tempfile

rep, obj, kern, out_md, title = sys.argv[1:6]
top = int(sys.argv[6]) if len(sys.argv) > 6 else 30
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
sass = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(sass) if ".text." in l and kern in l][0]
cur, inst = None, []
for l in sass[start + 1:]:
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l) and not l.strip().startswith("."):
        inst.append(cur)
    if (l.startswith(".section") or l.startswith("//----")) and len(inst) > 100:
        break
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr)]
assert abs(len(data) - len(inst)) < 64, (len(data), len(inst), "report and object do not match")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
by, why = collections.Counter(), collections.defaultdict(collections.Counter)
for i, r in enumerate(data[:len(inst)]):
    by[inst[i]] += int(r[idx["# Samples"]] or 0)
    for s in stalls:
        why[inst[i]][s] += int(r[idx[s]] or 0)
total = sum(by.values())
src = {}
with open(out_md, "w") as f:
    f.write(f"# {title}\n\nWarp-state samples ({total}) per CUDA source line: SASS of the ncu report matched with "
            f"`nvdisasm --print-line-info` of `{os.path.basename(obj)}`.\n\n| samples | share | line | main stall |\n|---:|---:|---|---|\n")
    for key, c in by.most_common(top):
        if key is None:
            continue
        fn, ln = key
        if fn not in src:
            p = [q for q in glob.glob(os.path.join(ROOT, "vision_assist_b200", "csrc", "*")) if q.endswith(fn)]
            src[fn] = open(p[0]).read().splitlines() if p else []
        text = src[fn][ln - 1].strip()[:100].replace("|", "\\|") if src[fn] and ln <= len(src[fn]) else ""
        st, n = why[key].most_common(1)[0]
        f.write(f"| {c} | {100 * c / total:.1f}% | `{fn}:{ln}` `{text}` | {st[6:]} ({n}) |\n")
print("wrote", out_md)
