set -x
timeout 600 python benchmarks/overlap_probe.py 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r2b_overlap.log
cat gpurun_out/r2b_overlap.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"k_search|k_mask|k_popcount" -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2b_ncu_launch.log 2>&1
tail -2 gpurun_out/r2b_ncu_launch.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_search|k_mask" --launch-skip 24 -c 8 -o gpurun_out/r2b_full -f python benchmarks/prof_step.py --steps 8 > gpurun_out/r2b_ncu_full.log 2>&1
tail -3 gpurun_out/r2b_ncu_full.log
ls -la gpurun_out/
