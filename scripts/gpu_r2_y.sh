( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 ) > gpurun_out/r2y_tests.log 2>&1
cat gpurun_out/r2y_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py --gpus 1 > gpurun_out/r2y_bench_n1.json 2> gpurun_out/r2y_bench_n1.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2y_bench_n1.json").read().strip().splitlines()[-1])
print("value %.1fM ms/step %.4f e2e %.1fM frac %.3f frac_traffic %.3f traffic %.1f MB" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["roofline"]["frac"], d["roofline"]["frac_of_traffic"], d["roofline"]["traffic"]/1e6))
oc=d["other_configs"]
print("cfg0 %.1f us/pair" % oc["configs0_sample_corpus_batch1"]["us_per_query_pair"], oc["configs0_sample_corpus_batch1"]["parity_vs_golden"])
print({k:(round(v["fused_kernel_ms"]*1e3,1),round(v["fused_kernel_frac_of_hbm"],3),v["parity_vs_oracle"]) for k,v in oc["configs3_fused_decode_step"].items() if isinstance(v,dict)})
print("1B: %.1fM e2e %.1fM" % (oc["configs2_1B_tokens_one_gpu"]["value"]/1e6, oc["configs2_1B_tokens_one_gpu"]["e2e_value"]/1e6), d["clocks"])
PY
