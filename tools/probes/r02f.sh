python -m pytest tests/test_rollout.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02f_rollout_pytest.log; cat gpurun_out/r02f_rollout_pytest.log
{
for mt in 1 2; do for rep in 1 2; do FOOTSIES_B200_ROLLOUT_MT=$mt python tools/rollout_sweep.py --one 16384 64; done; done
FOOTSIES_B200_ROLLOUT_FFMA=1 python tools/rollout_sweep.py --one 16384 64
for mt in 1 2; do FOOTSIES_B200_ROLLOUT_MT=$mt python tools/rollout_sweep.py --one 131072 64; FOOTSIES_B200_ROLLOUT_MT=$mt python tools/rollout_sweep.py --one 1048576 64; done
FOOTSIES_B200_ROLLOUT_FFMA=1 python tools/rollout_sweep.py --one 1048576 64
FOOTSIES_B200_ROLLOUT_MT=1 python tools/rollout_sweep.py --one 16384 32
python tools/rollout_sweep.py --self-play
} > gpurun_out/r02f_rollout_mma.log 2>&1
cat gpurun_out/r02f_rollout_mma.log
