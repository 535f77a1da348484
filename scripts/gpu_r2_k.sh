timeout 100 benchmarks/micro/comp_probe | tee gpurun_out/r2i_comp_probe.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_processors.py tests/test_soft_boost.py -m gpu -x -q -k "logits or mask or fused or processor or boost or decode" 2>&1 | tail -8
timeout 300 python -m pytest tests/test_gpu_robustness.py -m gpu -x -q 2>&1 | tail -8
timeout 200 python benchmarks/decode_probe.py 2>&1 | tail -9 | tee gpurun_out/r2k_decode_row.log
FMIDX_LOGITS_PATH=regs timeout 200 python benchmarks/decode_probe.py 2>&1 | tail -9 | tee gpurun_out/r2k_decode_regs.log
