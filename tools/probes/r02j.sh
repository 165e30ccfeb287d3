{
for thr in 0 100000000; do
  for args in "--envs 4096 --k 1 --graph" "--envs 65536 --k 4 --selfplay --graph" "--envs 65536 --k 1 --graph" "--envs 262144 --k 1 --graph" "--envs 524288 --k 1 --graph" "--envs 1048576 --k 1 --graph" "--envs 1048576 --k 4 --graph --steps 100"; do
    echo "== pdl_min_envs=$thr $args"
    FOOTSIES_B200_PDL_MIN_ENVS=$thr python tools/quick_bench.py --reps 2 --burnin 300 $args 2>&1 | tail -1
  done
done
} > gpurun_out/r02j_pdl_threshold.log 2>&1
cat gpurun_out/r02j_pdl_threshold.log
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
