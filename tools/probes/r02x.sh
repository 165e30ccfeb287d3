# r02x: plain-C ABI parity program on the B200; compute-sanitizer (memcheck / racecheck / synccheck) on that native binary and on
# the rollout workload (round 1 could not get the sanitizer to run on this pool: record what happens now)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_c_abi.py -m gpu -x -q 2>&1 | tail -3
./tests/c_abi/abi_vs_oracle > gpurun_out/r02x_c_abi_parity.log 2>&1; echo "c_abi rc=$?"; tail -6 gpurun_out/r02x_c_abi_parity.log
{
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool tests/c_abi/abi_vs_oracle 10"
  timeout 300 compute-sanitizer --tool $tool --print-limit 5 ./tests/c_abi/abi_vs_oracle 10 2>&1 | tail -12; echo "rc=${PIPESTATUS[0]}"
done
echo "== compute-sanitizer --tool racecheck python tools/sanitize_rollout.py"
SAN_HORIZON=6 timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitize_rollout.py 2>&1 | tail -12; echo "rc=${PIPESTATUS[0]}"
} > gpurun_out/r02x_sanitizer.log 2>&1
cat gpurun_out/r02x_sanitizer.log
