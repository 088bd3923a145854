#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (ncu --set full) as the metric,unit,value CSV kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep profiles/rNN_kernel_ncu_full.csv "header comment"
"""
import csv
import io
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "sm__cycles_active.avg",
        "sm__cycles_elapsed.avg", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct")


def main():
    rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, vals = rows[0], rows[1], rows[2]
    with open(out, "w") as f:
        f.write("# %s\nmetric,unit,value\n" % note)
        for n, u, v in zip(names, units, vals):
            if n == "Kernel Name" or n in KEEP or (n.startswith("smsp__average_warps_issue_stalled") and "not_issued" not in n):
                f.write("%s,%s,%s\n" % (n, u, v.replace(",", "")))
    print("wrote", out)


if __name__ == "__main__":
    main()
