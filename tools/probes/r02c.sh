python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02c_pytest.log
cat gpurun_out/r02c_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "rc=$?"
tail -c 3000 gpurun_out/r02c_bench.json; tail -5 gpurun_out/r02c_bench.err
./tools/probes/mma_probe > gpurun_out/r02_mma_probe.log 2>&1; cat gpurun_out/r02_mma_probe.log
