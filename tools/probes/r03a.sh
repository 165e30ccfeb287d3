# r03a: rollout tests incl. the 32-battle-warp shape (forced on small ragged batches, and selected by size), bench with the half-size lead slice
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout.py -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r03a.json 2> gpurun_out/bench_r03a.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r03a.json").read().strip().splitlines()[-1])
print("value %.4g frac %.4f e2e %.4g slowest %.4g natural %.4g" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["value_slowest_block"], d["e2e"]["natural_width_layout"]["value"]))
for k, v in d["extra"].items(): print(k, v.get("env_frames_per_sec"), v.get("e2e_env_frames_per_sec"))
PY
