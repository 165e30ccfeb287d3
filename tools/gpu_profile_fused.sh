#!/bin/bash
# ncu full captures of the fused-K step kernel (issue slots / occupancy / branch efficiency instead of GB/s):
# config C (65 536 envs self-play, K = 4) and 1 Mi envs vs bot, K = 4.  usage: tools/gpu_profile_fused.sh <tag>
TAG=${1:-x}
OUT=gpurun_out
C="python tools/quick_bench.py --burnin 300 --steps 20 --reps 1 --envs 65536 --k 4 --selfplay"
D="python tools/quick_bench.py --burnin 300 --steps 20 --reps 1 --envs 1048576 --k 4"
$C > $OUT/plain_fusedC_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 310 -c 2 -f -o $OUT/prof_fusedC_$TAG $C > $OUT/ncu_fusedC_$TAG.log 2>&1
$D > $OUT/plain_fusedD_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 310 -c 2 -f -o $OUT/prof_fusedD_$TAG $D > $OUT/ncu_fusedD_$TAG.log 2>&1
echo done
