# r02y: round-2 ncu captures of the fused-K step kernel (issue slots / occupancy / lanes), configs[2] and 1 Mi battles K = 4
bash tools/gpu_profile_fused.sh r02y
tail -1 gpurun_out/plain_fusedC_r02y.log gpurun_out/plain_fusedD_r02y.log
ls -la gpurun_out/*.ncu-rep
