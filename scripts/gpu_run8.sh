set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2h_tests.log
cat gpurun_out/r2h_tests.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err ) 2>&1 | tail -3
tail -12 gpurun_out/r2h_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2h_bench.json").read().strip().splitlines()[-1])
print("value %.1fM ms/step %.4f e2e %.1fM launches %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["gpu_launches"]))
r=d["roofline"]; print("roofline achieved %.0f frac %.3f" % (r["achieved"], r["frac"]), r["other_kernel_ms"], r["counts_per_step"])
print(json.dumps(d["ablation"])[:300])
print(json.dumps(d["cpu_baseline"])[:300])
PY
