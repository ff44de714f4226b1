"""Summarise an .ncu-rep (raw page) into a small text table for profiles/ (run where ncu is installed).

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
    "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_not_selected",
    "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_no_instructions",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("=" * 100)
        print(r[ki])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("  {:72s} {:>18s} {}".format(w, r[i], units[i]))
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(r[hdr.index("dram__bytes_write.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_write.sum")]]
            t = float(r[hdr.index("gpu__time_duration.sum")]) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}[units[hdr.index("gpu__time_duration.sum")]]
            print("  {:72s} {:>18.1f} MB per launch".format("traffic = dram read + write", (rd + wr) / 1e6))
            print("  {:72s} {:>18.1f} GB/s (under ncu, cold cache)".format("dram bandwidth", (rd + wr) / t / 1e9))
        except Exception:
            pass


if __name__ == "__main__":
    main(sys.argv[1])
