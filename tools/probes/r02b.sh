set -x
python -m pytest tests/test_gpu_actor_modes.py tests/test_gpu_parity.py tests/test_battle_state.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02b_pytest.log
cat gpurun_out/r02b_pytest.log
for rep in 1 2; do
for v in base pdl rev rng all; do
  echo "== $v eager"; FOOTSIES_B200_LIB=$PWD/tools/probes/lib_$v.so python tools/quick_bench.py --reps 2 --burnin 400 --steps 400 2>&1 | tail -1
done; done > gpurun_out/r02b_ab.log 2>&1
for v in base all; do
  echo "== $v graph"; FOOTSIES_B200_LIB=$PWD/tools/probes/lib_$v.so python tools/quick_bench.py --reps 2 --burnin 400 --steps 200 --graph 2>&1 | tail -1
  for e in 1048576 2097152 8388608; do echo "== $v envs $e"; FOOTSIES_B200_LIB=$PWD/tools/probes/lib_$v.so python tools/quick_bench.py --reps 2 --burnin 400 --steps 200 --envs $e 2>&1 | tail -1; done
done >> gpurun_out/r02b_ab.log 2>&1
cat gpurun_out/r02b_ab.log
