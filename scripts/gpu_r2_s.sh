for cfg in "24 0" "32 0" "24 1"; do
set -- $cfg
echo "== BENCH_NBUF=$1 FMIDX_SHARD_PROBE=$2"
FMIDX_PULL_TRACE=1 BENCH_NBUF=$1 FMIDX_SHARD_PROBE=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --mode sharded > gpurun_out/r2s_sharded4_$1_$2.json 2> gpurun_out/r2s_sharded4_$1_$2.err
grep "rank 0 doc_sharded\|rank 0.: \|pull trace (rank 0)\|rror" gpurun_out/r2s_sharded4_$1_$2.err | tail -3
done
