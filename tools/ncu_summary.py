#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU) into the per-kernel numbers DESIGN.md / profiles/ quote.
usage: tools/ncu_summary.py report.ncu-rep [more metric substrings...]"""
import csv, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "lts__t_bytes.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sector_hit_rate.pct"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    cols = [h for h in hdr if h in WANT or any(e in h for e in extra)]
    for r in data:
        print("== " + r[ki][:110])
        for c in cols:
            i = hdr.index(c)
            print("   %-75s %s %s" % (c, r[i], units[i]))


if __name__ == "__main__":
    main()
