( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 ) > gpurun_out/r2_final_tests_2gpu.log 2>&1
cat gpurun_out/r2_final_tests_2gpu.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r2_final_bench_n2.json 2> gpurun_out/r2_final_bench_n2.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_final_bench_n2.json").read().strip().splitlines()[-1])
print("value %.1fM ms/step %.4f e2e %.1fM launches %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["gpu_launches"]))
for k,v in d["modes"].items(): print(k, {a:b for a,b in v.items() if a in ("value","e2e_value","ms_per_step","parity_vs_bruteforce_on_sample","buffer_sets")} if isinstance(v, dict) else v)
PY
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_final_ref_n2.json 2> gpurun_out/r2_final_ref_n2.err ) 2>&1 | tail -3
cut -c1-160 gpurun_out/r2_final_ref_n2.json
