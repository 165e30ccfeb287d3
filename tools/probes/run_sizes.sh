#!/bin/bash
for f in "$@"; do
  echo "== $f"
  for args in "--envs 524288 --k 1 --graph" "--envs 262144 --k 4 --graph --steps 100" "--envs 65536 --k 4 --selfplay --graph" "--envs 65536 --k 1 --graph" "--envs 16384 --k 1 --graph" "--envs 4096 --k 1 --graph"; do
    echo "-- $args"
    FOOTSIES_B200_LIB=$PWD/$f python tools/quick_bench.py --reps 2 --burnin 300 $args 2>&1 | tail -1
  done
done
