( timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) > gpurun_out/r2q_tests.log 2>&1
cat gpurun_out/r2q_tests.log
( time timeout 900 python bench.py --gpus 1 > gpurun_out/r2q_bench_n1.json 2> gpurun_out/r2q_bench_n1.err ) 2>&1 | tail -3
tail -8 gpurun_out/r2q_bench_n1.err | cut -c1-400
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2q_bench_n1.json").read().strip().splitlines()[-1])
print("value %.1fM ms/step %.4f e2e %.1fM (eager %.1fM) launches %s frac %.3f traffic %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["eager_api"]["value"]/1e6, d["gpu_launches"], d["roofline"]["frac"], d["roofline"]["traffic"]))
print(d["other_configs"]["configs0_sample_corpus_batch1"])
print({k:v for k,v in d["other_configs"]["configs2_1B_tokens_one_gpu"].items() if k in ("value","e2e_value")})
PY
