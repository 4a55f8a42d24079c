#!/bin/bash
# ncu --set full of the kernels of the split blending path (unfiltered states, 65,536 robots, split forced on)
tag=${1:-r02_split_prof}
export SAI_B200_BLEND_SPLIT=1
ncu --set full --clock-control none --import-source on -k regex:osc_blend_ -s 60 -c 3 -o gpurun_out/${tag}_full -f \
    python bench.py --steps 6 --warmup 3 --min-ratio 0 --no-cpu > gpurun_out/${tag}_ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_full.ncu-rep gpurun_out/${tag}_ncu_selected.csv osc_blend > gpurun_out/${tag}_ncu_selected.txt 2>&1
python tools/ncu_source_lines.py gpurun_out/${tag}_full.ncu-rep osc_blend_variants 40 > gpurun_out/${tag}_variants_lines.txt 2>&1
python tools/ncu_source_lines.py gpurun_out/${tag}_full.ncu-rep osc_blend_classify 25 > gpurun_out/${tag}_classify_lines.txt 2>&1
head -50 gpurun_out/${tag}_variants_lines.txt
