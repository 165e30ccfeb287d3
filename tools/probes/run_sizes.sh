#!/bin/bash
# usage: run_sizes.sh lib1 lib2 ...  -> quick_bench (CUDA-graph replay) at several sizes per lib
for f in "$@"; do
  echo "== $f"
  for args in "--envs 1048576 --k 4 --steps 100" "--envs 262144 --k 1" "--envs 65536 --k 4 --selfplay" "--envs 65536 --k 1" "--envs 16384 --k 1" "--envs 4096 --k 1"; do
    echo "-- $args"
    FOOTSIES_B200_LIB=$PWD/$f python tools/quick_bench.py --graph --reps 2 --burnin 300 $args 2>&1 | tail -1
  done
done
