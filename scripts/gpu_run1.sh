set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2a_tests.log
for occ in 2 3 4 5; do
  FMIDX_BYTES_OCC=$occ timeout 300 python bench.py --steps 2000 --no-cpu-baseline > gpurun_out/r2a_occ$occ.json 2> gpurun_out/r2a_occ$occ.err
done
FMIDX_BYTES_OCC=3 BENCH_NPIPE=6 CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python bench.py --steps 2000 --no-cpu-baseline > gpurun_out/r2a_occ3_p6.json 2> gpurun_out/r2a_occ3_p6.err
FMIDX_BYTES_OCC=2 BENCH_NPIPE=8 CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python bench.py --steps 2000 --no-cpu-baseline > gpurun_out/r2a_occ2_p8.json 2> gpurun_out/r2a_occ2_p8.err
cat gpurun_out/r2a_tests.log
for f in gpurun_out/r2a_occ*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"]/1e6, d["ms_per_step"], d["pipeline"]["serial_ms_per_step"], d["roofline"]["other_kernel_ms"], d["e2e"]["value"]/1e6)
except Exception as e: print("ERR", e)
PY
done
