#!/bin/bash
# BASELINE config 5: batch-size sweep of config 2 on one GPU, closing state of round 2 (FP64, FP32 mode, unfiltered states)
out=gpurun_out/r02_sweep_final.jsonl
: > $out
for r in 1024 4096 16384 65536 262144 1048576 4194304; do
  sets=8; [ $r -ge 1048576 ] && sets=2
  python bench.py --robots $r --steps 200 --warmup 5 --sets $sets --no-cpu 2>/dev/null | tail -1 >> $out
done
python - <<PY
import json
print("| robots | cycles/s | roofline fraction | p99 latency ms | end to end | unfiltered states | FP32 mode |")
print("|---:|---:|---:|---:|---:|---:|---:|")
for l in open("$out"):
    d = json.loads(l)
    f = d["extra"].get("fp32_mode")
    print("| %s | %.3g | %.3f | %.4f | %.3g | %.3g | %s |" % (format(d["config"]["robots_per_gpu"], ","), d["value"], d["roofline"]["frac"], d["latency_ms"]["p99"], d["e2e"]["value"],
          d["extra"]["unfiltered_states"]["value"], ("%.3g" % f["value"]) if f else "-"))
PY
