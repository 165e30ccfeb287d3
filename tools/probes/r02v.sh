# r02v: (1) ragged multi-chunk batches through every CTA shape (the reverse-traversal fix), (2) bulk-store variant: correctness
# then interleaved A/B against the shipped library, (3) the library from before the fix on the same ragged batch (expected: hangs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
{
echo "== bulk-store variant, tools/sanitize_check.py (vs oracle), three shape / PDL settings"
for cfg in "0 0" "1000000000 0" "0 1000000000"; do set -- $cfg
  FOOTSIES_B200_LIB=$PWD/tools/probes/lib_bulk.so FOOTSIES_B200_LARGE_SHAPE_MIN_ENVS=$1 FOOTSIES_B200_PDL_MIN_ENVS=$2 SAN_STEPS=120 timeout 300 python tools/sanitize_check.py 2>&1 | tail -3
done
echo "== A/B 4 Mi battles"
bash tools/probes/run_ab.sh "--envs 4194304" footsies_gym_b200/libfootsies_b200.so tools/probes/lib_bulk.so
echo "== A/B 1 Mi battles (CUDA graph)"
bash tools/probes/run_ab.sh "--envs 1048576 --graph" footsies_gym_b200/libfootsies_b200.so tools/probes/lib_bulk.so
echo "== A/B 65536 battles (CUDA graph)"
bash tools/probes/run_ab.sh "--envs 65536 --graph" footsies_gym_b200/libfootsies_b200.so tools/probes/lib_bulk.so
} > gpurun_out/r02v_bulk_store_ab.log 2>&1
cat gpurun_out/r02v_bulk_store_ab.log
echo "== library from before the ragged-tail fix, 230 000 battles (second launch walks backwards)"
FOOTSIES_B200_LIB=$PWD/tools/probes/lib_before_ragged_fix.so SAN_STEPS=4 timeout -s KILL 60 python tools/sanitize_check.py > gpurun_out/r02v_before_fix.log 2>&1; echo "rc=$?" >> gpurun_out/r02v_before_fix.log
tail -5 gpurun_out/r02v_before_fix.log
nvidia-smi --query-gpu=name,utilization.gpu --format=csv
