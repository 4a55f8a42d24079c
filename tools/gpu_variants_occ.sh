#!/bin/bash
# occupancy experiment on the variants kernel of the split blending path (registers capped for 4 / 6 / 8 blocks per SM)
tag=${1:-r02_occ}
SAI_B200_BLEND_SPLIT=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${tag}_gputest_eager.log; tail -3 gpurun_out/${tag}_gputest_eager.log
for minb in 4 6 8; do
  export SAI_B200_VARIANTS_MINB=$minb
  for r in 65536 262144 1048576; do
    sets=8; [ $r -ge 262144 ] && sets=4; [ $r -ge 1048576 ] && sets=2
    SAI_B200_BLEND_SPLIT=1 python bench.py --robots $r --sets $sets --steps 100 --warmup 10 --min-ratio 0 --no-cpu 2>/dev/null > gpurun_out/${tag}_unfiltered_minb${minb}_$r.json
    python -c "import json;d=json.loads(open('gpurun_out/${tag}_unfiltered_minb${minb}_$r.json').read().strip().splitlines()[-1]);print('minb',$minb,$r,'%.4g cycles/s'%d['value'],'%.4f ms'%d['ms_per_step'])"
  done
done
