( time timeout 900 python bench.py --gpus 1 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err ) 2>&1 | tail -3
tail -3 gpurun_out/r2_final_bench_n1.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_final_bench_n1.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.1fM ms/step %.4f e2e %.1fM" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6))
print("frac %.3f frac_traffic %.3f sector frac %.3f" % (r["frac"], r["frac_of_traffic"], r["sector_rate"]["frac_of_random_gather"]))
for k,v in d["other_configs"]["configs3_fused_decode_step"].items():
    if isinstance(v, dict): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ("ms_per_step","fused_kernel_ms","fused_kernel_frac_of_hbm","ms_per_step_incremental","incremental_matches_full_search","parity_vs_oracle")})
PY
