python -m pytest tests/test_rollout.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -4
{
for rep in 1 2; do python tools/rollout_sweep.py --one 16384 64; done
python tools/rollout_sweep.py --one 16384 32
python tools/rollout_sweep.py --one 131072 64
python tools/rollout_sweep.py --one 1048576 64
python tools/rollout_sweep.py --self-play
} > gpurun_out/r02l_rollout_trunc_split.log 2>&1
cat gpurun_out/r02l_rollout_trunc_split.log
