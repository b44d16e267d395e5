#!/usr/bin/env python3
"""tools/ncu_summary.py <report.ncu-rep> [substring...] -- print selected raw metrics per launch
(read here, on the CPU box, from reports brought back in gpurun_out/)."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__occupancy_limit_blocks", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.max"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("--- %s  grid=%s block=%s" % (r[hdr.index("Kernel Name")][:70], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for i, h in enumerate(hdr):
            base = h.split(".", 2)[-1] if h.count(".") > 2 and h.split(".")[1] in ("TriageCompute",) else h
            if h in KEYS or any(x in h for x in extra):
                print("  %-78s %s %s" % (h, r[i], units[i]))


if __name__ == "__main__":
    main()
