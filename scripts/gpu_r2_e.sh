timeout 200 python benchmarks/overlap_probe2.py 2>&1 | grep -v "^$" | tail -8 | tee gpurun_out/r2e_probe2.log
FMIDX_K1_ZERO=1 timeout 200 python benchmarks/overlap_probe2.py 2>&1 | grep -v "^$" | tail -8 | tee -a gpurun_out/r2e_probe2.log
