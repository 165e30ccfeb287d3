# r02z: short first slice of the host-buffer path (A/B by FOOTSIES_B200_HOST_LEAD_DIV), host-path GPU tests with the new slice plan
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_actor_modes.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python tools/e2e_bench.py --lead > gpurun_out/r02z_e2e_lead_slice.log 2>&1; cat gpurun_out/r02z_e2e_lead_slice.log
