{
for mt in 1 2; do FOOTSIES_B200_ROLLOUT_MT=$mt python tools/rollout_sweep.py --one 131072 64; FOOTSIES_B200_ROLLOUT_MT=$mt python tools/rollout_sweep.py --one 1048576 64; FOOTSIES_B200_ROLLOUT_MT=$mt python tools/rollout_sweep.py --one 65536 64; done
} > gpurun_out/r02m_rollout_mt.log 2>&1
cat gpurun_out/r02m_rollout_mt.log
