#!/bin/bash
# Split blending path (classification / variants / fallback kernels) against the single blending kernel: parity tests both
# ways, unfiltered-state throughput at several batch sizes (the host picks the path from the hand-over count), launch list.
tag=${1:-r02_split}
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${tag}_gputest_lazy.log; tail -3 gpurun_out/${tag}_gputest_lazy.log
SAI_B200_BLEND_SPLIT=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${tag}_gputest_eager.log; tail -3 gpurun_out/${tag}_gputest_eager.log
for mode in auto single; do
  unset SAI_B200_BLEND_SPLIT
  [ $mode = single ] && export SAI_B200_BLEND_SPLIT=0
  [ $mode = split ] && export SAI_B200_BLEND_SPLIT=1
  for r in 65536 131072 262144 1048576 4194304; do
    sets=8; [ $r -ge 262144 ] && sets=4; [ $r -ge 1048576 ] && sets=2; [ $r -ge 4194304 ] && sets=1
    python bench.py --robots $r --sets $sets --steps 100 --warmup 10 --min-ratio 0 --no-cpu 2>/dev/null > gpurun_out/${tag}_unfiltered_${mode}_$r.json
    python -c "import json;d=json.loads(open('gpurun_out/${tag}_unfiltered_${mode}_$r.json').read().strip().splitlines()[-1]);print('$mode',$r,'%.4g cycles/s'%d['value'],'%.4f ms'%d['ms_per_step'],'launches',d.get('gpu_launches'))"
  done
done
unset SAI_B200_BLEND_SPLIT
python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null > gpurun_out/${tag}_bench_20.json
python -c "import json;d=json.loads(open('gpurun_out/${tag}_bench_20.json').read().strip().splitlines()[-1]);print('filtered','%.4g'%d['value'],d['roofline']['frac'],d['latency_ms']['p50'],d['extra'].get('unfiltered_states',{}).get('value'))"
SAI_B200_BLEND_SPLIT=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_unfiltered.csv \
    python bench.py --steps 6 --warmup 3 --min-ratio 0 --no-cpu > gpurun_out/${tag}_ncu.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${tag}_launches_unfiltered.csv")) if len(r) > 10 and r[0].isdigit()]
d = collections.defaultdict(list)
for r in rows:
    d[r[4][:70]].append(float(r[-1]))
for k, v in d.items():
    if "osc_" not in k: continue
    v2 = sorted(v)
    print(k, "n=%d median=%.2f us min=%.2f max=%.2f" % (len(v), v2[len(v2)//2] / 1e3, v2[0] / 1e3, v2[-1] / 1e3))
PY
