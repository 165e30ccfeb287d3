N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02p_bench_${N}gpu.json 2> gpurun_out/r02p_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02p_bench_${N}gpu.json"))
print("N=%d value %.4e frac %.4f e2e %.3e natural %.3e" % (d["n_gpus"], d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["natural_width_layout"]["value"]))
print({k: "%.3e" % v["env_frames_per_sec"] for k,v in d["extra"].items()})
PY
