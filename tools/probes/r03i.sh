# r03i: does a one-off stagger between the two CTAs of an SM help the rollout kernel at 16 384 battles?
mkdir -p gpurun_out
{
for rep in 1 2; do
for lib in footsies_gym_b200/libfootsies_b200.so tools/probes/lib_stag800.so tools/probes/lib_stag1600.so; do
echo "== $lib"
FOOTSIES_B200_LIB=$PWD/$lib python tools/rollout_sweep.py --one 16384 64
done; done
} > gpurun_out/r03i_rollout_stagger.log 2>&1
cat gpurun_out/r03i_rollout_stagger.log
