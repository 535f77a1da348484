timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -15
for g in 1 0; do
echo "== BENCH_SHARD_GRAPH=$g"
BENCH_SHARD_GRAPH=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode sharded --big-tokens 400000000 > gpurun_out/r2p_sharded_graph$g.json 2> gpurun_out/r2p_sharded_graph$g.err
grep "doc_sharded\|rror" gpurun_out/r2p_sharded_graph$g.err | tail -3
done
