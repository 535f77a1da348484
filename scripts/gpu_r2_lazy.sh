timeout 600 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
( time timeout 900 python bench.py --gpus 1 --no-other-configs > gpurun_out/r2_lazy_bench_n1.json 2> gpurun_out/r2_lazy_bench_n1.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_lazy_bench_n1.json").read().strip().splitlines()[-1])
print("value %.1fM ms/step %.4f e2e %.1fM (eager %.1fM)" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["eager_api"]["value"]/1e6))
print("full zeroing: %.1fM" % (d["ablation"]["full_row_zeroing"]["value"]/1e6), "table off %.1fM" % (d["ablation"]["unigram_table_off"]["value"]/1e6))
print(d["cpu_baseline"]["parity_vs_gpu_on_sample"], d["e2e"]["matches_device_path"])
PY
