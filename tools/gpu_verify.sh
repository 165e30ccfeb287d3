#!/bin/bash
# Verification pass on one GPU box: every GPU test, smoke(), both bench arms.  usage: tools/gpu_verify.sh <tag>
set -u
TAG=${1:-verify}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "rc=$?" >> $OUT/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "rc=$?" >> $OUT/smoke_$TAG.log
# both arms as the driver calls them
python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "rc=$?" >> $OUT/bench_$TAG.err
python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
echo done
