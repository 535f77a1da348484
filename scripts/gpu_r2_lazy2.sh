M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,smsp__inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum
timeout 280 ncu --replay-mode range --clock-control none --metrics $M --csv --log-file gpurun_out/r2_range_replay_lazy.csv python benchmarks/range_probe.py --phase 0 --lazy 2 > gpurun_out/r2_range_replay_lazy.log 2>&1
tail -1 gpurun_out/r2_range_replay_lazy.log; grep "dram__bytes\|inst_executed" gpurun_out/r2_range_replay_lazy.csv | cut -d, -f11-13
cp gpurun_out/r2_range_replay_lazy.csv profiles/r2_range_replay_lazy.csv
( time timeout 900 python bench.py --gpus 1 > gpurun_out/r2_lazy_bench_n1.json 2> gpurun_out/r2_lazy_bench_n1.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_lazy_bench_n1.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.1fM ms/step %.4f e2e %.1fM (eager %.1fM)" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["eager_api"]["value"]/1e6))
print("roofline: alg %.1f MB traffic %s frac %.3f frac_traffic %s sector_rate %s" % (r["algorithmic_bytes_per_launch"]/1e6, r["traffic"], r["frac"], r["frac_of_traffic"], r["sector_rate"]))
print("full zeroing:", d["ablation"]["full_row_zeroing"])
oc=d["other_configs"]
print("1B: %.1fM e2e %.1fM" % (oc["configs2_1B_tokens_one_gpu"]["value"]/1e6, oc["configs2_1B_tokens_one_gpu"]["e2e_value"]/1e6), oc["configs2_1B_tokens_one_gpu"]["parity_vs_bruteforce_on_sample"])
print("cfg0 %.1f" % oc["configs0_sample_corpus_batch1"]["us_per_query_pair"], d["cpu_baseline"]["parity_vs_gpu_on_sample"], d["e2e"]["matches_device_path"])
PY
