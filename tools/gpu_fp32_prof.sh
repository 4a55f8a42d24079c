#!/bin/bash
# ncu --set full of the single-precision instantiation of the fused kernel (bench.py's extra.fp32_mode section)
tag=${1:-r02_fp32}
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:osc32 -s 20 -c 2 -o gpurun_out/${tag}_full -f \
    python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_full.ncu-rep gpurun_out/${tag}_ncu_selected.csv osc_cycle_kernel > gpurun_out/${tag}_ncu_selected.txt 2>&1
ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for k in ['Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__icc_request_hit_rate.pct','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','dram__bytes_read.sum','dram__bytes_write.sum']:
    if k in h: print(k, [r[h.index(k)][:60] for r in rows[2:]])
"
