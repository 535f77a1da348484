set -x
timeout 600 python benchmarks/decode_probe.py 2>&1 | tail -12
FMIDX_LOGITS_PATH=tma timeout 600 python benchmarks/decode_probe.py 2>&1 | tail -12
