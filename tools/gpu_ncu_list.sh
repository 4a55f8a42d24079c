#!/bin/bash
# launch list (per-kernel durations, serialised by ncu) of a short bench run, with and without cross-cycle pipelining
tag=${1:-r02}
for mode in pipe nopipe; do
  if [ $mode = nopipe ]; then export SAI_B200_NO_PIPELINE=1; else unset SAI_B200_NO_PIPELINE; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_${mode}.csv \
      python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_${mode}.log 2>&1
  python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${tag}_launches_${mode}.csv")) if len(r) > 10 and r[0].isdigit()]
d = collections.defaultdict(list)
for r in rows:
    d[r[4][:60]].append(float(r[-1]))
for k, v in d.items():
    v2 = sorted(v)
    print("${mode}", k, "n=%d median=%.2f us min=%.2f max=%.2f" % (len(v), v2[len(v2)//2] / 1e3, v2[0] / 1e3, v2[-1] / 1e3))
PY
done
