#!/bin/bash
# interleaved A/B runs in ONE call (boxes differ by a few % from call to call): usage run_ab.sh "<quick_bench args>" lib...
ARGS=$1; shift
for rep in 1 2 3; do
  for f in "$@"; do
    echo "== $f"
    FOOTSIES_B200_LIB=$PWD/$f python tools/quick_bench.py --reps 1 --burnin 300 $ARGS 2>&1 | tail -1
  done
done
