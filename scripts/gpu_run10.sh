set -x
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 600 python benchmarks/decode_probe.py 2>&1 | tail -9
