set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2e_tests.log
cat gpurun_out/r2e_tests.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err ) 2>&1 | tail -3
tail -12 gpurun_out/r2e_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2e_bench.json").read().strip().splitlines()[-1])
print("value %.1fM ms/step %.4f e2e %.1fM launches %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["gpu_launches"]))
r=d["roofline"]; print("roofline achieved %.0f frac %.3f" % (r["achieved"], r["frac"]), r["other_kernel_ms"], r["counts_per_step"])
print(json.dumps(d["ablation"])[:400])
print(json.dumps(d["timing"]))
PY
for occ in 3 4 6; do FMIDX_BYTES_OCC=$occ timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('occ $occ value %.1fM e2e %.1fM serial %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['other_kernel_ms']))"; done
BENCH_NPIPE=6 CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('npipe6 value %.1fM e2e %.1fM' % (d['value']/1e6, d['e2e']['value']/1e6))"
BENCH_NPIPE=8 CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('npipe8 value %.1fM e2e %.1fM' % (d['value']/1e6, d['e2e']['value']/1e6))"
