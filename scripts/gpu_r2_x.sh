timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_processors.py tests/test_reference_golden_gpu.py -m gpu -x -q -k "logits or fused or processor or golden or decode" 2>&1 | tail -3
timeout 200 python benchmarks/decode_probe.py 2>&1 | tail -8 | tee gpurun_out/r2x_decode.log
