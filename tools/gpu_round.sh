#!/bin/bash
# One GPU session: parity tests, bench (both arms), ncu launch list and one full capture of the fused kernel.
# usage (through gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_$tag.log
tail -3 $out/pytest_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
cat $out/bench_$tag.json
python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "ref rc=$?"
cat $out/bench_ref_$tag.json
python bench.py --steps 6 --warmup 3 --no-cpu > $out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu > $out/ncu_launch_$tag.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:osc_cycle_kernel -s 4 -c 2 -f -o $out/prof_$tag \
    python bench.py --steps 6 --warmup 3 --no-cpu > $out/ncu_full_$tag.log 2>&1
echo "ncu full rc=$?"
