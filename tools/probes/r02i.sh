python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02i_pytest.log; cat gpurun_out/r02i_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02i_bench_ref.json 2> gpurun_out/r02i_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r02i_bench_ref.json
python examples/ppo_footsies.py > gpurun_out/r02i_ppo_example.log 2>&1; tail -5 gpurun_out/r02i_ppo_example.log
