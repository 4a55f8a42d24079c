#!/bin/bash
# ncu --set full of the rolled general-path kernel on config 3 with uniformly sampled states (13.4 k of 262,144 robots)
tag=${1:-r02_generic}
ncu --set full --clock-control none --import-source on -k regex:osc_singular_kernel -s 5 -c 1 -o gpurun_out/${tag}_full -f \
    python tools/bench_configs.py 4096 262144 > gpurun_out/${tag}_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_full.ncu-rep gpurun_out/${tag}_ncu_selected.csv osc_singular > gpurun_out/${tag}_ncu_selected.txt 2>&1
python tools/ncu_source_lines.py gpurun_out/${tag}_full.ncu-rep osc_singular_kernel 30 > gpurun_out/${tag}_lines.txt 2>&1
head -45 gpurun_out/${tag}_lines.txt
grep -E "gpu__time_duration|inst_executed.sum|issue_active|local_ld|local_st|l1tex__t_sector_hit|stalled_(long|wait|no_inst|short)" gpurun_out/${tag}_ncu_selected.txt | head -12
