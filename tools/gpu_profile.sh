#!/bin/bash
# Run on the GPU box through gpurun: bench, then the ncu launch list and one full capture of the step kernel.
# usage: tools/gpu_profile.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
BENCH_ARGS="--steps 20 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 2"
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?" >> $OUT/bench_$TAG.err
python bench.py $BENCH_ARGS > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py $BENCH_ARGS > $OUT/ncu_launches_$TAG.log 2>&1
python bench.py $BENCH_ARGS > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 8 -c 2 -f -o $OUT/prof_step_$TAG \
    python bench.py $BENCH_ARGS > $OUT/ncu_full_$TAG.log 2>&1
echo done
