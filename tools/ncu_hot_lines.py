"""Hottest SASS lines of a kernel from `ncu -i X.ncu-rep --page source --csv`: stall samples per instruction with the
dominant stall reasons, in program order around the hot spots.

  python tools/ncu_hot_lines.py source.csv [top]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ci, si = hdr.index("Source"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "not_issued" not in h.lower()]
recs = []
for k, r in enumerate(rows[2:]):
    try:
        s = int(r[si])
    except (ValueError, IndexError):
        continue
    reasons = []
    for i, h in stall_cols:
        try:
            v = int(r[i])
        except (ValueError, IndexError):
            v = 0
        if v:
            reasons.append((v, h.replace("stall_", "")))
    reasons.sort(reverse=True)
    recs.append((k, s, r[ci], reasons[:3]))
tot = sum(r[1] for r in recs)
print("instructions", len(recs), "samples", tot)
agg = {}
for _, s, _, reasons in recs:
    for v, h in reasons:
        agg[h] = agg.get(h, 0) + v
print("stall totals (top-3 per line only):", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for k, s, src, reasons in sorted(recs, key=lambda r: -r[1])[:top]:
    print("{:5d} {:6d} {:5.2f}%  {:70s} {}".format(k, s, 100.0 * s / max(tot, 1), src[:70], " ".join("%s=%d" % (h, v) for v, h in reasons)))
