#!/bin/bash
# GPU-box check used during development: parity tests, then the bench the way the driver runs it, with and without
# cross-cycle pipelining, then the default (1000-step) bench.  Results under gpurun_out/<tag>_*.
tag=${1:-r02}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_gputest.log; tail -5 gpurun_out/${tag}_gputest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_20.json 2> gpurun_out/${tag}_bench_20.err
SAI_B200_NO_PIPELINE=1 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_20_nopipe.json 2>/dev/null
python bench.py --no-cpu > gpurun_out/${tag}_bench_1000.json 2>/dev/null
for f in 20 20_nopipe 1000; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.4g frac %.3f" % (d["value"], d["roofline"]["frac"]), d["timing"]["brackets"], "lat", d["latency_ms"]["p50"], d["latency_ms"]["p99"],
          "e2e %.4g" % d["e2e"]["value"], d["e2e"]["host_link"]["gbs"], d["e2e"]["host_link"]["e2e_fraction_of_link"], d["extra"]["device_resident_one_stream_per_instance"]["value"], d.get("cpu_baseline"))
except Exception as e:
    print("$f failed", e)
PY
done
tail -5 gpurun_out/${tag}_bench_20.err
