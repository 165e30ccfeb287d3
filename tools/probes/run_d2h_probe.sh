#!/bin/bash
# Aggregate pinned-copy bandwidth of the box: N = 1, 2, 4, 8 processes (one GPU each) copying at the same time.
#   tools/probes/run_d2h_probe.sh <max gpus> > profiles/r02_d2h_probe.log
# Variants: default vs NUMA-local pinned allocation; one staged block vs six pieces per iteration; d2h, h2d and both.
set -u
cd "$(dirname "$0")"
nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o d2h_probe d2h_probe.cu || exit 1
MAX=${1:-8}
echo "# host: $(nproc) cpus, numa nodes: $(ls -d /sys/devices/system/node/node* 2>/dev/null | wc -l), $(free -g | awk '/Mem:/{print $2}') GiB"
nvidia-smi topo -m 2>/dev/null | head -14
for n in 1 2 4 8; do
  [ "$n" -gt "$MAX" ] && break
  variants=("113 1 0 d2h" "113 1 1 d2h")
  [ "$n" -eq "$MAX" ] && variants+=("113 6 1 d2h" "64 1 1 d2h" "16 1 1 d2h" "113 1 1 h2d" "113 1 1 both")
  for variant in "${variants[@]}"; do
    start=$(python3 -c 'import time; print(time.time() + 1.5)')
    echo "== processes=$n block_mib/pieces/numa/dir = $variant"
    for ((g = 0; g < n; g++)); do ./d2h_probe $g 2.0 $variant $start & done | sort | tee /tmp/d2h_$$.log
    wait
    python3 - /tmp/d2h_$$.log <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
print("   aggregate %.1f GB/s over %d processes (min %.1f, max %.1f per process)" % (
    sum(r["gbps"] for r in rows), len(rows), min(r["gbps"] for r in rows), max(r["gbps"] for r in rows)))
PY
  done
done
