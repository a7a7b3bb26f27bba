#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, on the CPU box) into the counters DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [substring filters...]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
]


def main():
    rep = sys.argv[1]
    filt = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("== %s  grid=%s block=%s" % (name[:90], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for i, h in enumerate(hdr):
            short = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1].startswith("Triage") else h
            hit = any(h.endswith(k) or short == k for k in KEYS) or "warp_issue_stalled" in h and "per_warp_active" in h
            if filt:
                hit = hit or any(f in h for f in filt)
            if hit and r[i] not in ("", "0", "n/a"):
                print("   %-95s %s %s" % (h[-95:], r[i], units[i]))


if __name__ == "__main__":
    main()
