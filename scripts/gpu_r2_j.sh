timeout 100 benchmarks/micro/comp_probe | tee gpurun_out/r2i_comp_probe.log
for pr in 0 1 2; do
echo "== FMIDX_SHARD_PROBE=$pr"
FMIDX_SHARD_PROBE=$pr timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode sharded --big-tokens 400000000 > gpurun_out/r2j_sharded_probe$pr.json 2> gpurun_out/r2j_sharded_probe$pr.err
grep "doc_sharded" gpurun_out/r2j_sharded_probe$pr.err | tail -2
done
