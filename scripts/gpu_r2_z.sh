timeout 300 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_parity.py tests/test_reference_golden_gpu.py -m gpu -x -q -k "scalar or golden or sample or surface or edge" 2>&1 | tail -4
timeout 300 python - <<'PY'
import json, os, sys, time
sys.path.insert(0, ".")
import bench
for spec in ("1", "0"):
    os.environ["FMIDX_SCALAR_SPECULATE"] = spec
    r = bench.other_config0(0)
    print("speculate", spec, "%.1f us/pair" % r["us_per_query_pair"], r["parity_vs_golden"])
PY
