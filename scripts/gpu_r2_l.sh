timeout 300 python -m pytest tests/test_gpu_robustness.py tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -6
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused" 2>&1 | tail -3
for nb in 8 16; do
echo "== BENCH_NBUF=$nb"
BENCH_NBUF=$nb timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode sharded --big-tokens 400000000 > gpurun_out/r2l_sharded_nbuf$nb.json 2> gpurun_out/r2l_sharded_nbuf$nb.err
grep "doc_sharded" gpurun_out/r2l_sharded_nbuf$nb.err | tail -1
done
for pr in 1 2; do
echo "== FMIDX_SHARD_PROBE=$pr"
FMIDX_SHARD_PROBE=$pr timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode sharded --big-tokens 400000000 > gpurun_out/r2l_sharded_probe$pr.json 2> gpurun_out/r2l_sharded_probe$pr.err
grep "doc_sharded" gpurun_out/r2l_sharded_probe$pr.err | tail -1
done
echo "== FMIDX_PULL_CTAS=296"
FMIDX_PULL_CTAS=296 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode sharded --big-tokens 400000000 > gpurun_out/r2l_sharded_pull296.json 2> gpurun_out/r2l_sharded_pull296.err
grep "doc_sharded" gpurun_out/r2l_sharded_pull296.err | tail -1
