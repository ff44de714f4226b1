"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X) per kernel.

  python tools/launch_summary.py gpurun_out/launches.csv "<command that was profiled>" > profiles/<name>_summary.txt
"""
import csv
import re
import sys
from collections import OrderedDict


def main(path, cmd):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
    agg = OrderedDict()
    for r in rd:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").strip()
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ui], 1e-3)
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + v * scale)
    # input generation and torch fills are not part of the step
    agg = OrderedDict((k, v) for k, v in agg.items() if not k.startswith(("synth_pv_kernel", "at::")))
    total = sum(t for _, t in agg.values())
    print("ncu launch list of `{}` (gpu__time_duration.sum, --clock-control none)".format(cmd))
    print("cold-cache, serialised launches: compare SHARES with bench.py's kernel_ms_share, not absolutes\n")
    for name, (n, t) in agg.items():
        print("{:45s} launches={:3d} total={:8.3f} ms avg={:9.1f} us share={:.4f}".format(name, n, t / 1e3, t / n, t / total))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "?")
