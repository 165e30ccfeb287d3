python tools/rollout_sweep.py --one 1048576 64 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rollout_mma_kernel -s 2 -c 1 -f -o gpurun_out/prof_rollout_1Mi_r02r python tools/rollout_sweep.py --one 1048576 64 > gpurun_out/ncu_rollout_1Mi_r02r.log 2>&1
tail -2 gpurun_out/ncu_rollout_1Mi_r02r.log
