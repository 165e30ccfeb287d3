#!/bin/bash
# Round-1 (h) profile pass on one GPU box: both bench arms, the ncu launch list and full captures of the step kernel
# (steady state) and of the whole-horizon rollout kernel; each ncu pass only after the same command exited 0 without ncu.
set -u
TAG=${1:-r01h}
OUT=gpurun_out
mkdir -p $OUT
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?" >> $OUT/bench_$TAG.err
python bench.py --impl reference --steps 200 --warmup 5 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
python bench.py --workload rollout --steps 50 --warmup 3 > $OUT/bench_rollout_$TAG.json 2> $OUT/bench_rollout_$TAG.err
SHORT="--steps 20 --warmup 3 --burnin 0 --no-cpu-baseline --no-extra --e2e-steps 2"
python bench.py $SHORT > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py $SHORT > $OUT/ncu_launches_$TAG.log 2>&1
STEADY="--steps 20 --warmup 3 --burnin 600 --no-cpu-baseline --no-extra --e2e-steps 2"
python bench.py $STEADY > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 610 -c 2 -f -o $OUT/prof_step_$TAG \
    python bench.py $STEADY > $OUT/ncu_full_$TAG.log 2>&1
python tools/rollout_sweep.py --one 16384 64 > $OUT/rollout_plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -f -o $OUT/prof_rollout_$TAG \
    python tools/rollout_sweep.py --one 16384 64 > $OUT/ncu_rollout_$TAG.log 2>&1
python tools/e2e_bench.py > $OUT/e2e_$TAG.log 2>&1
python examples/ppo_footsies.py > $OUT/ppo_example_$TAG.log 2>&1
echo done
