# r03e: layer 2 of the policy with the output tile as the outer loop (lib_ntouter.so, -DFG_POLICY_L2_NT_OUTER=1) against the shipped order
mkdir -p gpurun_out
{
FOOTSIES_B200_LIB=$PWD/tools/probes/lib_ntouter.so python -m pytest tests/test_rollout.py -m gpu -x -q 2>&1 | tail -2
for rep in 1 2 3; do
for lib in footsies_gym_b200/libfootsies_b200.so tools/probes/lib_ntouter.so; do
echo "== $lib"
FOOTSIES_B200_LIB=$PWD/$lib python tools/rollout_sweep.py --one 16384 64
done; done
for lib in footsies_gym_b200/libfootsies_b200.so tools/probes/lib_ntouter.so; do
echo "== $lib"
FOOTSIES_B200_LIB=$PWD/$lib python tools/rollout_sweep.py --one 131072 64
FOOTSIES_B200_LIB=$PWD/$lib python tools/rollout_sweep.py --one 1048576 64
FOOTSIES_B200_LIB=$PWD/$lib python tools/rollout_sweep.py --one 16384 32
done
} > gpurun_out/r03e_rollout_l2_order.log 2>&1
cat gpurun_out/r03e_rollout_l2_order.log
