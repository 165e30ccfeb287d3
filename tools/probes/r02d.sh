for e in 1 2; do for rep in 1 2; do FOOTSIES_B200_ROLLOUT_E=$e python tools/rollout_sweep.py --one 16384 64; done; done > gpurun_out/r02d_rollout_e.log 2>&1
for e in 1 2 4; do FOOTSIES_B200_ROLLOUT_E=$e python tools/rollout_sweep.py --one 131072 64; done >> gpurun_out/r02d_rollout_e.log 2>&1
FOOTSIES_B200_ROLLOUT_E=1 python -m pytest tests/test_rollout.py -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/r02d_rollout_e.log
cat gpurun_out/r02d_rollout_e.log
