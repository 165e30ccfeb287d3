# r03c (N GPUs): bench under torchrun with the final library + the D2H probe of this box.  usage: r03c.sh N
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
$TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r03c_${N}gpu.json 2> gpurun_out/bench_r03c_${N}gpu.err; echo "bench rc=$?"
python - $N <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench_r03c_{sys.argv[1]}gpu.json").read().strip().splitlines()[-1])
print("N=%d value %.4g frac %.4f e2e %.4g natural %.4g" % (d["n_gpus"], d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["natural_width_layout"]["value"]))
for k, v in d["extra"].items(): print(k, v.get("env_frames_per_sec"), v.get("e2e_env_frames_per_sec"))
PY
nproc; numactl -H 2>/dev/null | head -3; lscpu | grep -i "numa\|model name\|socket" | head -6
