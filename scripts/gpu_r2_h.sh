for b in 1024 16384 65536; do
echo "== batch $b"
timeout 200 python benchmarks/overlap_probe.py --batch $b --pipes 2,4,8,16 --steps 200 2>&1 | grep "both\|graph" | tee -a gpurun_out/r2h_batch_sweep.log
done
