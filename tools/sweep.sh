#!/bin/bash
# BASELINE config 5: batch-size sweep of config 2 on one GPU (run through gpurun): bash tools/sweep.sh <tag>
tag=${1:-r01}
out=gpurun_out/sweep_$tag.jsonl
: > $out
for r in 1024 4096 16384 65536 262144 1048576; do
  sets=8; [ $r -ge 262144 ] && sets=4; [ $r -ge 1048576 ] && sets=2
  python bench.py --robots $r --sets $sets --steps 200 --warmup 10 --no-cpu 2>/dev/null >> $out
done
python - "$out" <<'PY'
import json, sys
print("| robots | cycles/s (one stream) | ms per cycle | fraction of FP64 roofline | cycles/s, one stream per instance | end to end (host buffers) | p99 cycle latency ms |")
print("|---:|---:|---:|---:|---:|---:|---:|")
for l in open(sys.argv[1]):
    d = json.loads(l)
    print("| %d | %.3g | %.4f | %.3f | %.3g | %.3g | %.4f |" % (d["config"]["robots_per_gpu"], d["value"], d["ms_per_step"], d["roofline"]["frac"],
          d["extra"]["device_resident_one_stream_per_instance"]["value"], d["e2e"]["value"], d["latency_ms"]["p99"]))
PY
