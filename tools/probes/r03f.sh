# r03f: 32-battle-warp rollout kernel compiled for 3 CTAs per SM (168 registers, 24 B of spill) against the shipped 190 registers
mkdir -p gpurun_out
{
FOOTSIES_B200_LIB=$PWD/tools/probes/lib_mt2b3.so python -m pytest tests/test_rollout.py -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do
for lib in footsies_gym_b200/libfootsies_b200.so tools/probes/lib_mt2b3.so; do
echo "== $lib"
FOOTSIES_B200_LIB=$PWD/$lib python tools/rollout_sweep.py --one 131072 64
FOOTSIES_B200_LIB=$PWD/$lib python tools/rollout_sweep.py --one 1048576 64
done; done
echo "== shipped library, 16 384 battles"
python tools/rollout_sweep.py --one 16384 64
} > gpurun_out/r03f_rollout_mt2_blocks.log 2>&1
cat gpurun_out/r03f_rollout_mt2_blocks.log
