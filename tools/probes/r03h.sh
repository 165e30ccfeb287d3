# r03h: 8-battle warps (half M tiles, FOOTSIES_B200_ROLLOUT_MT=0) for small rollout batches against the shipped 16-battle warps
mkdir -p gpurun_out
{
python -m pytest tests/test_rollout.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2 3; do
for mt in 1 0; do
echo "== MT=$mt"
FOOTSIES_B200_ROLLOUT_MT=$mt python tools/rollout_sweep.py --one 16384 64
done; done
for n in 4096 8192 32768 65536; do for mt in 1 0; do echo "== MT=$mt"; FOOTSIES_B200_ROLLOUT_MT=$mt python tools/rollout_sweep.py --one $n 64; done; done
for mt in 1 0; do echo "== MT=$mt H=32"; FOOTSIES_B200_ROLLOUT_MT=$mt python tools/rollout_sweep.py --one 16384 32; done
} > gpurun_out/r03h_rollout_half_tiles.log 2>&1
cat gpurun_out/r03h_rollout_half_tiles.log
