timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"k_search|k_mask|k_popcount" -c 400 --csv --log-file gpurun_out/r2_launches_lazy.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2_ncu_launch_lazy.log 2>&1
tail -1 gpurun_out/r2_ncu_launch_lazy.log | cut -c1-120
