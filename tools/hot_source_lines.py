"""Stall samples of a kernel by CUDA source line: joins the SASS view of an ncu source page
(`ncu -i X.ncu-rep --page source --csv`, instruction order) with nvdisasm's line info of the same cubin.

  python tools/hot_source_lines.py <source.csv> <cubin> <kernel-substring> [top]
"""
import collections
import csv
import re
import subprocess
import sys

csv_path, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, inside = [], None, False
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
si = hdr.index("# Samples")
stall_cols = [(i, h.replace("stall_", "")) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
samples = []
for r in rows[2:]:
    try:
        samples.append((int(r[si]), r))
    except (ValueError, IndexError):
        pass
print("sass instructions: ncu", len(samples), "nvdisasm", len(lines))
agg, why = collections.Counter(), collections.defaultdict(collections.Counter)
for k, (s, r) in enumerate(samples):
    key = lines[k] if k < len(lines) and lines[k] else ("?", 0)
    agg[key] += s
    for i, h in stall_cols:
        try:
            v = int(r[i])
        except (ValueError, IndexError):
            v = 0
        if v:
            why[key][h] += v
tot = sum(agg.values())
src_cache = {}
for (f, l), s in agg.most_common(top):
    text = ""
    try:
        if f not in src_cache:
            src_cache[f] = open("wavebreaking_b200/csrc/" + f).read().splitlines()
        text = src_cache[f][l - 1].strip()[:90]
    except Exception:
        pass
    print("{:5.1f}%  {}:{:<5d} {:92s} {}".format(100.0 * s / max(tot, 1), f, l, text,
                                              " ".join("%s=%d" % kv for kv in why[(f, l)].most_common(3))))
