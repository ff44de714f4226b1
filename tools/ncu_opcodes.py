"""Executed warp instructions by opcode from `ncu -i X.ncu-rep --page source --csv` (which SASS dominates a kernel)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci, ei, si = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
cnt, smp, tot, stot = collections.Counter(), collections.Counter(), 0, 0
for r in rows[2:]:
    try:
        n, s = int(r[ei]), int(r[si])
    except (ValueError, IndexError):
        continue
    parts = r[ci].split()
    op = parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "?")
    op = op.split(".")[0]
    cnt[op] += n
    smp[op] += s
    tot += n
    stot += s
print("total warp instructions", tot, "samples", stot)
for k, v in cnt.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    print("{:12s} {:12d} {:5.1f}%   samples {:5.1f}%".format(k, v, 100.0 * v / tot, 100.0 * smp[k] / max(stot, 1)))
