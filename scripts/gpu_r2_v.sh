( time timeout 900 python bench.py --gpus 1 > gpurun_out/r2v_bench_n1.json 2> gpurun_out/r2v_bench_n1.err ) 2>&1 | tail -3
grep "^value" gpurun_out/r2v_bench_n1.err
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2v_ref_n1.json 2> gpurun_out/r2v_ref_n1.err ) 2>&1 | tail -3
cut -c1-200 gpurun_out/r2v_ref_n1.json
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,smsp__inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum
timeout 280 ncu --replay-mode range --clock-control none --metrics $M --csv --log-file gpurun_out/r2v_range_phase0.csv python benchmarks/range_probe.py --phase 0 > gpurun_out/r2v_range_phase0.log 2>&1
tail -1 gpurun_out/r2v_range_phase0.log; grep "dram__bytes" gpurun_out/r2v_range_phase0.csv | cut -d, -f11-13
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"k_search|k_mask|k_popcount" -c 400 --csv --log-file gpurun_out/r2v_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2v_ncu_launch.log 2>&1
tail -1 gpurun_out/r2v_ncu_launch.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_search|k_mask" --launch-skip 24 -c 8 -o gpurun_out/r2v_full -f python benchmarks/prof_step.py --steps 8 > gpurun_out/r2v_ncu_full.log 2>&1
tail -2 gpurun_out/r2v_ncu_full.log
