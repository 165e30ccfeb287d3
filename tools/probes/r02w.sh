# r02w: is the lower e2e of the no-flag bench run (4000 launches per block) the long burn before it, or the box?  Same box, back to back.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-extra > gpurun_out/r02w_bench_steps20_a.json 2>/dev/null
python bench.py --no-extra > gpurun_out/r02w_bench_default.json 2>/dev/null
python bench.py --steps 20 --warmup 5 --no-extra > gpurun_out/r02w_bench_steps20_b.json 2>/dev/null
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02w_bench_reference_arm.json 2>/dev/null
python - <<'PY'
import json
for f in ("steps20_a", "default", "steps20_b"):
    d = json.loads(open(f"gpurun_out/r02w_bench_{f}.json").read().strip().splitlines()[-1])
    print(f, d["steps"], d["blocks"], "value %.4g" % d["value"], "frac %.4f" % d["roofline"]["frac"], "e2e %.4g" % d["e2e"]["value"],
          "slowest %.4g" % d["e2e"]["value_slowest_block"], "natural %.4g" % d["e2e"]["natural_width_layout"]["value"], d["clocks"]["reasons"], d["clocks"]["power_w_max"])
d = json.loads(open("gpurun_out/r02w_bench_reference_arm.json").read().strip().splitlines()[-1])
print("reference arm", d["value"], d["steps"], d["cpu_baseline"]["sample"][:100])
PY
python tools/fuzz_kernel_vs_oracle.py --backend gpu --master-seed 5 1.5 > gpurun_out/r02w_fuzz_gpu_vs_oracle.log 2>&1; tail -1 gpurun_out/r02w_fuzz_gpu_vs_oracle.log; grep -c "steps=60" gpurun_out/r02w_fuzz_gpu_vs_oracle.log
