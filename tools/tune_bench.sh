#!/bin/bash
# bench every tuning build (run through gpurun): bash tools/tune_bench.sh [bench args]
for lib in sai_primitives_b200/_tune/lib_*.so; do
  n=$(basename $lib .so)
  SAI_B200_OSC_LIB=$PWD/$lib python bench.py --steps 100 --warmup 10 --no-cpu "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-28s value %.4g  kernel_ms %.4f  e2e %.4g' % ('$n', d['value'], d['roofline']['kernel_ms'], d['e2e']['value']))"
done
