python -m pytest tests/test_rollout.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02g_pytest.log; cat gpurun_out/r02g_pytest.log
{
for w in 4 7; do for rep in 1 2; do FOOTSIES_B200_ROLLOUT_W=$w python tools/rollout_sweep.py --one 16384 64; done; done
python tools/rollout_sweep.py --one 16384 64
python tools/rollout_sweep.py --one 16384 32
python tools/rollout_sweep.py --one 65536 64
python tools/rollout_sweep.py --self-play
} > gpurun_out/r02g_rollout_mma.log 2>&1
cat gpurun_out/r02g_rollout_mma.log
python tools/rollout_sweep.py --one 16384 64 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rollout_mma_kernel -s 2 -c 1 -f -o gpurun_out/prof_rollout_r02g python tools/rollout_sweep.py --one 16384 64 > gpurun_out/ncu_rollout_r02g.log 2>&1
tail -2 gpurun_out/ncu_rollout_r02g.log
