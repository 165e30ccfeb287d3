#!/bin/bash
# Round-2 profile pass on one GPU box (B200_PROFILING.md recipe): every ncu pass only after the same command exited 0
# without ncu; numbers printed under ncu are never bench values.
#   (1) launch list of a short bench run                      -> gpurun_out/launches_$TAG.csv
#   (2) ncu --set full on the steady-state step kernel         -> gpurun_out/prof_step_$TAG.ncu-rep  (kernel replay: cold L2)
#   (3) DRAM traffic of the same launches in their NATURAL cache state (application replay, no cache control): the step
#       kernel alternates its traversal direction per launch to reuse what the previous launch left in L2, which kernel
#       replay (flush + restore between passes) cannot see                     -> gpurun_out/traffic_$TAG.csv
#   (4) ncu --set full on the whole-horizon rollout kernel     -> gpurun_out/prof_rollout_$TAG.ncu-rep
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
SHORT="--steps 20 --warmup 3 --burnin 0 --no-cpu-baseline --no-extra --e2e-steps 2"
python bench.py $SHORT > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py $SHORT > $OUT/ncu_launches_$TAG.log 2>&1
STEADY="--steps 20 --warmup 3 --burnin 600 --no-cpu-baseline --no-extra --e2e-steps 2"
python bench.py $STEADY > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 700 -c 2 -f -o $OUT/prof_step_$TAG \
    python bench.py $STEADY > $OUT/ncu_full_$TAG.log 2>&1
ncu --replay-mode application --cache-control none --clock-control none \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
    -k regex:step_kernel -s 700 -c 4 --csv --log-file $OUT/traffic_$TAG.csv \
    python bench.py $STEADY > $OUT/ncu_traffic_$TAG.log 2>&1
python tools/rollout_sweep.py --one 16384 64 > $OUT/rollout_plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rollout_ -s 2 -c 1 -f -o $OUT/prof_rollout_$TAG \
    python tools/rollout_sweep.py --one 16384 64 > $OUT/ncu_rollout_$TAG.log 2>&1
echo done
