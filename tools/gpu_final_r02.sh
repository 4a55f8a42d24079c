#!/bin/bash
# round-2 closing check on one GPU: smoke, the driver's bench lines (both arms), the default bench
out=gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
SECONDS=0; python bench.py --steps 20 --warmup 3 > $out/r02_final_bench_20.json 2> $out/r02_final_bench_20.err; echo "bench --steps 20 wall $SECONDS s"; tail -1 $out/r02_final_bench_20.err
SECONDS=0; python bench.py --impl reference --steps 20 --warmup 3 > $out/r02_final_bench_ref.json 2> $out/r02_final_bench_ref.err; echo "reference arm wall $SECONDS s"; tail -1 $out/r02_final_bench_ref.err
SECONDS=0; python bench.py --no-cpu > $out/r02_final_bench_1000.json 2> $out/r02_final_bench_1000.err; echo "default bench wall $SECONDS s"; tail -1 $out/r02_final_bench_1000.err
python - <<PY
import json
for f in ("20", "1000"):
    d = json.loads(open("gpurun_out/r02_final_bench_%s.json" % f).read().strip().splitlines()[-1])
    print(f, "value %.4g frac %.4f" % (d["value"], d["roofline"]["frac"]), "lat", d["latency_ms"]["p50"], "e2e %.4g" % d["e2e"]["value"], "launches", d["gpu_launches"],
          "unfiltered %.4g" % d["extra"]["unfiltered_states"]["value"], "1M %.4g" % d["extra"]["unfiltered_states_1m_robots"]["value"], (d.get("cpu_baseline") or {}).get("value"))
d = json.loads(open("gpurun_out/r02_final_bench_ref.json").read().strip().splitlines()[-1])
print("reference arm", d["value"], d["unit"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"])
PY
