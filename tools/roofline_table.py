"""Per-kernel roofline table from a tools/ncu_summary.py text file (one `--set full` capture of one batch).

  python tools/roofline_table.py profiles/r1m_ncu_full_all_kernels.txt 148 > profiles/r1m_kernel_roofline.md
"""
import json
import re
import sys

UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def parse(path):
    kernels, cur = [], None
    for line in open(path):
        if line.startswith("===="):
            cur = None
            continue
        if cur is None:
            cur = {"name": line.strip()}
            kernels.append(cur)
            continue
        m = re.match(r"\s+(\S+)\s+([-\d.,e+]+)\s*(\S*)", line)
        if m:
            try:
                cur[m.group(1)] = (float(m.group(2).replace(",", "")), m.group(3))
            except ValueError:
                pass
    return kernels


def main(path, steps):
    try:
        peak = json.load(open("MEASURED_PEAKS.json"))
        hbm = float(peak.get("hbm_gbs", peak.get("hbm_copy_gbs", 6553.0)))
    except Exception:
        hbm = 6553.0
    seen = set()
    print("| kernel | time (us) | us / step | DRAM read+write (MB) | DRAM GB/s | % of measured HBM peak ({:.0f} GB/s) | issue active % | FP64 pipe % | warps active % | regs | top stall |".format(hbm))
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for k in parse(path):
        name = re.sub(r"\(.*", "", k["name"]).replace("void ", "")
        if name in seen or name.startswith(("synth_pv", "at::")) or "gpu__time_duration.sum" not in k:
            continue
        seen.add(name)
        t = k["gpu__time_duration.sum"][0] * UNIT.get(k["gpu__time_duration.sum"][1], 1.0)
        rd = k["dram__bytes_read.sum"][0] * UNIT[k["dram__bytes_read.sum"][1]]
        wr = k["dram__bytes_write.sum"][0] * UNIT[k["dram__bytes_write.sum"][1]]
        gbs = (rd + wr) / (t * 1e-6) / 1e9
        stalls = {s.replace("smsp__pcsamp_warps_issue_stalled_", ""): v[0] for s, v in k.items() if isinstance(v, tuple) and "issue_stalled" in s and "selected" not in s}
        top = max(stalls, key=stalls.get) if stalls else "-"
        print("| `{}` | {:.1f} | {:.2f} | {:.1f} | {:.0f} | {:.1f} | {:.0f} | {:.0f} | {:.0f} | {:.0f} | {} |".format(
            name, t, t / steps, (rd + wr) / 1e6, gbs, 100.0 * gbs / hbm,
            k["smsp__issue_active.avg.pct_of_peak_sustained_active"][0],
            k["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0],
            k["sm__warps_active.avg.pct_of_peak_sustained_active"][0], k["launch__registers_per_thread"][0], top))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 148)
