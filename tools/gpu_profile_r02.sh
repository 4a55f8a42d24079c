#!/bin/bash
# round-2 profile artefacts: full bench line (driver style), ncu --set full of the fused kernel, batch-size sweep (config 5)
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_final20.json 2> gpurun_out/r02_bench_final20.err || tail -5 gpurun_out/r02_bench_final20.err
ncu --set full --clock-control none --import-source on -k regex:osc_cycle_kernel -s 30 -c 3 -o gpurun_out/r02_cycle_full -f \
    python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/r02_ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_cycle_full.ncu-rep gpurun_out/r02_ncu_selected.csv osc_cycle_kernel > gpurun_out/r02_ncu_selected.txt 2>&1
tail -45 gpurun_out/r02_ncu_selected.txt
for r in 1024 4096 16384 65536 262144 1048576 4194304; do
  sets=8; [ $r -ge 1048576 ] && sets=2
  python bench.py --robots $r --steps 200 --warmup 5 --sets $sets --no-cpu 2>/dev/null | tail -1 >> gpurun_out/r02_sweep.jsonl
done
python - <<PY
import json
for l in open("gpurun_out/r02_sweep.jsonl"):
    d=json.loads(l); print(d["config"]["robots_per_gpu"], "%.4g" % d["value"], "%.3f" % d["roofline"]["frac"], "e2e %.4g" % d["e2e"]["value"], "unfiltered %.4g" % d["extra"]["unfiltered_states"]["value"])
PY
