timeout 100 python benchmarks/graph_debug.py 2>&1 | tail -12
timeout 240 python benchmarks/overlap_probe.py --pipes 4,8,16 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r2c_overlap.log
cat gpurun_out/r2c_overlap.log
