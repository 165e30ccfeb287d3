#!/bin/bash
# Run on the GPU box through gpurun: GPU tests, both bench arms, then the ncu launch list and one full capture of the
# step kernel in steady state (after the 600-step burn-in).  Every ncu pass runs only after the same command exited 0
# without ncu.  usage: tools/gpu_profile.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "rc=$?" >> $OUT/pytest_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?" >> $OUT/bench_$TAG.err
python bench.py --impl reference --steps 200 --warmup 5 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
SHORT="--steps 20 --warmup 3 --burnin 0 --no-cpu-baseline --no-extra --e2e-steps 2"
python bench.py $SHORT > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py $SHORT > $OUT/ncu_launches_$TAG.log 2>&1
STEADY="--steps 20 --warmup 3 --burnin 600 --no-cpu-baseline --no-extra --e2e-steps 2"
python bench.py $STEADY > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 610 -c 2 -f -o $OUT/prof_step_$TAG \
    python bench.py $STEADY > $OUT/ncu_full_$TAG.log 2>&1
echo done
