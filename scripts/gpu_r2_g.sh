M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum
for ph in 0 1 2; do
timeout 280 ncu --replay-mode range --clock-control none --metrics $M --csv --log-file gpurun_out/r2g_range_phase$ph.csv python benchmarks/range_probe.py --phase $ph > gpurun_out/r2g_range_phase$ph.log 2>&1
tail -2 gpurun_out/r2g_range_phase$ph.log
cat gpurun_out/r2g_range_phase$ph.csv | tail -20
done
