#!/usr/bin/env python
"""Selected counters of an .ncu-rep -> CSV (kept under profiles/).  usage: ncu_summary.py rep out.csv [kernel-substring]"""
import csv, subprocess, sys

KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
    "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
    "smsp__inst_executed_op_branch.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
]
STALL = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
KEYS += [STALL % s for s in ("no_instruction", "long_scoreboard", "wait", "short_scoreboard", "math_pipe_throttle",
                             "barrier", "branch_resolving", "not_selected", "selected", "mio_throttle", "lg_throttle")]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    sub = sys.argv[3] if len(sys.argv) > 3 else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEYS if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            if sub in r[hdr.index("Kernel Name")]:
                w.writerow([r[i] for i in idx])
    for r in rows[2:]:
        if sub in r[hdr.index("Kernel Name")]:
            for i in idx:
                print("%-75s %s %s" % (hdr[i], r[i], units[i]))
            break


if __name__ == "__main__":
    main()
