#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): ``python tools/ncu_summary.py rep [pattern ...]``."""
import csv
import subprocess
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
           "lts__t_bytes.sum", "lts__t_sector_hit_rate", "sm__throughput.avg.pct", "sm__warps_active.avg.pct",
           "launch__registers_per_thread", "launch__occupancy_limit", "launch__grid_size", "launch__block_size",
           "sm__inst_executed.sum", "smsp__inst_executed.sum ", "sm__inst_executed_pipe_", "smsp__issue_active.avg.pct",
           "issue_stalled", "bank_conflicts", "wavefronts_mem_shared", "sm__cycles_elapsed.avg ", "sm__cycles_active.avg",
           "l1tex__t_bytes", "dynamic_shared", "sm__maximum_warps", "achieved_occupancy", "smsp__cycles_active.avg"]


def main() -> None:
    rep = sys.argv[1]
    pats = sys.argv[2:] or DEFAULT
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("kernels:", [r[hdr.index("Kernel Name")][:60] for r in data])
    for i, h in enumerate(hdr):
        if any(p in h for p in pats):
            vals = [r[i] for r in data]
            if all(v in ("0", "", "n/a") for v in vals) and "issue_stalled" in h:
                continue
            print(f"{h[:110]:110s} {units[i]:14s} {vals}")


if __name__ == "__main__":
    main()
