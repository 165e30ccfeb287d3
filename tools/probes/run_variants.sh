#!/bin/bash
for f in tools/probes/lib_*.so; do
  echo "== $f"
  FOOTSIES_B200_LIB=$PWD/$f python tools/quick_bench.py --reps 2 2>&1 | tail -1
done
