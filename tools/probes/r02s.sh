# FP16-split layers 2/3 of the rollout policy against the TF32 build (tools/probes/lib_tf32.so = -DFG_POLICY_F16=0)
python -m pytest tests/test_rollout.py -m gpu -x -q 2>&1 | tail -4
{
for rep in 1 2; do
echo "== f16 (default)"; python tools/rollout_sweep.py --one 16384 64
echo "== tf32"; FOOTSIES_B200_LIB=tools/probes/lib_tf32.so python tools/rollout_sweep.py --one 16384 64
done
for n in 131072 1048576; do
echo "== f16"; python tools/rollout_sweep.py --one $n 64
echo "== f16 MT=1"; FOOTSIES_B200_ROLLOUT_MT=1 python tools/rollout_sweep.py --one $n 64
echo "== tf32"; FOOTSIES_B200_LIB=tools/probes/lib_tf32.so python tools/rollout_sweep.py --one $n 64
done
echo "== f16 H=32"; python tools/rollout_sweep.py --one 16384 32
echo "== f16 self-play"; python tools/rollout_sweep.py --self-play
} > gpurun_out/r02s_rollout_f16.log 2>&1
cat gpurun_out/r02s_rollout_f16.log
