bash tools/probes/run_d2h_probe.sh 8 > gpurun_out/r02_d2h_probe.log 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)|Thread|Core" >> gpurun_out/r02_d2h_probe.log
tail -50 gpurun_out/r02_d2h_probe.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02h_bench_8gpu.json 2> gpurun_out/r02h_bench_8gpu.err
tail -c 2500 gpurun_out/r02h_bench_8gpu.json; tail -3 gpurun_out/r02h_bench_8gpu.err
