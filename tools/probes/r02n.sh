python -m pytest tests/test_rollout.py -m gpu -x -q 2>&1 | tail -2
{
for mt in 1 2; do FOOTSIES_B200_ROLLOUT_MT=$mt python tools/rollout_sweep.py --one 16384 64; FOOTSIES_B200_ROLLOUT_MT=$mt python tools/rollout_sweep.py --one 65536 64; FOOTSIES_B200_ROLLOUT_MT=$mt python tools/rollout_sweep.py --one 131072 64; FOOTSIES_B200_ROLLOUT_MT=$mt python tools/rollout_sweep.py --one 1048576 64; done
python tools/rollout_sweep.py --one 16384 32
} > gpurun_out/r02n_rollout_halves.log 2>&1
cat gpurun_out/r02n_rollout_halves.log
