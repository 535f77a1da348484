set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2b_tests.log
cat gpurun_out/r2b_tests.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 2000 --no-cpu-baseline > gpurun_out/r2b_$name.json 2> gpurun_out/r2b_$name.err; }
run base FMIDX_BYTES_OCC=5
run occ3 FMIDX_BYTES_OCC=3
run nok1z FMIDX_BYTES_OCC=5 FMIDX_K1_ZERO=0
run scan64k FMIDX_BYTES_OCC=5 FMIDX_SCAN_MAX=65536
run scan1k FMIDX_BYTES_OCC=5 FMIDX_SCAN_MAX=1024
run sampled FMIDX_BYTES_OCC=5 FMIDX_LAYOUT=sampled
for f in gpurun_out/r2b_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.1fM  ms/step %.4f serial %.4f  k: %s  e2e %.1fM" % (d["value"]/1e6, d["ms_per_step"], d["pipeline"]["serial_ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["other_kernel_ms"].items()}, d["e2e"]["value"]/1e6))
except Exception as e: print("ERR", e)
PY
done
