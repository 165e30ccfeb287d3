#!/bin/bash
# ncu full capture of the step kernel in steady state (after a 600-step burn-in). usage: tools/gpu_profile_steady.sh <tag> [extra quick_bench args]
TAG=${1:-x}; shift
OUT=gpurun_out
CMD="python tools/quick_bench.py --burnin 600 --steps 30 --reps 1 $@"
$CMD > $OUT/plain_steady_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 610 -c 2 -f -o $OUT/prof_steady_$TAG $CMD > $OUT/ncu_steady_$TAG.log 2>&1
echo done
