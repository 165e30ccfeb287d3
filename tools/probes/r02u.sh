# r02u: bench as the driver calls it, GPU differential fuzz (CUDA path vs oracle and vs oracle/_ref), ncu of the FP16 rollout kernel
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02u.json 2> gpurun_out/bench_r02u.err; echo "bench rc=$?"
python tools/fuzz_kernel_vs_oracle.py --backend gpu 2.5 > gpurun_out/r02u_fuzz_gpu_vs_oracle.log 2>&1; echo "fuzz rc=$?"; tail -1 gpurun_out/r02u_fuzz_gpu_vs_oracle.log
python tools/fuzz_kernel_vs_oracle.py --backend gpu --engine ref --master-seed 77 1.0 > gpurun_out/r02u_fuzz_gpu_vs_ref.log 2>&1; echo "fuzz ref rc=$?"; tail -1 gpurun_out/r02u_fuzz_gpu_vs_ref.log
python tools/rollout_sweep.py --one 16384 64 > gpurun_out/r02u_rollout.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rollout_mma_kernel -s 2 -c 1 -f -o gpurun_out/prof_rollout_r02u python tools/rollout_sweep.py --one 16384 64 > gpurun_out/ncu_rollout_r02u.log 2>&1
tail -2 gpurun_out/ncu_rollout_r02u.log; cat gpurun_out/r02u_rollout.log
