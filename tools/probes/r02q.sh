python -m pytest tests/test_rollout.py -m gpu -x -q 2>&1 | tail -2
{
for rep in 1 2 3; do python tools/rollout_sweep.py --one 16384 64; done
python tools/rollout_sweep.py --one 131072 64
python tools/rollout_sweep.py --one 1048576 64
} > gpurun_out/r02q_rollout.log 2>&1
cat gpurun_out/r02q_rollout.log
