# r03b (2 GPUs): bench under torchrun with the final library, reference arm under torchrun, the PPO example with gradients all-reduced
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r03b_2gpu.json 2> gpurun_out/bench_r03b_2gpu.err; echo "bench rc=$?"
$TR bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/bench_r03b_2gpu_ref.json 2> gpurun_out/bench_r03b_2gpu_ref.err; echo "ref rc=$?"
$TR examples/ppo_footsies.py --iters 12 > gpurun_out/r03b_ppo_2gpu.log 2>&1; echo "ppo rc=$?"; tail -4 gpurun_out/r03b_ppo_2gpu.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r03b_2gpu.json").read().strip().splitlines()[-1])
print("N=%d value %.4g frac %.4f e2e %.4g" % (d["n_gpus"], d["value"], d["roofline"]["frac"], d["e2e"]["value"]))
for k, v in d["extra"].items(): print(k, v.get("env_frames_per_sec"), v.get("e2e_env_frames_per_sec"))
r = json.loads(open("gpurun_out/bench_r03b_2gpu_ref.json").read().strip().splitlines()[-1])
print("reference arm lines:", len(open("gpurun_out/bench_r03b_2gpu_ref.json").read().strip().splitlines()), r["value"], r["steps"])
PY
