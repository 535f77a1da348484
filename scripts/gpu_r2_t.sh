timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -4
echo "== N=2 sharded 400M"
FMIDX_PULL_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode sharded --big-tokens 400000000 > gpurun_out/r2t_sharded2.json 2> gpurun_out/r2t_sharded2.err
grep "rank 0 doc_sharded\|pull trace (rank 0)\|rror" gpurun_out/r2t_sharded2.err | tail -3
