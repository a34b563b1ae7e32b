#!/usr/bin/env python
"""Summarise an ncu report (read here, on the CPU box): python tools/ncu_summary.py file.ncu-rep [metric-substring ...]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for r in data:
        print("### " + r[hdr.index("Kernel Name")])
        for i, h in enumerate(hdr):
            base = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1].startswith("Triage") else h
            if any(base == k or h == k for k in KEYS) or any(e in h for e in extra):
                print(f"  {h} = {r[i]} {units[i]}")


if __name__ == "__main__":
    main()
