timeout 250 python benchmarks/knob_sweep.py 2>&1 | tail -9 | tee gpurun_out/r2w_knobs.log
for occ in 2 3 6; do echo "FMIDX_BYTES_OCC=$occ"; FMIDX_BYTES_OCC=$occ timeout 200 python benchmarks/overlap_probe.py --pipes 8 --steps 300 2>&1 | grep "graph replay  " | tee -a gpurun_out/r2w_knobs.log; done
