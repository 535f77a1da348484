for cfg in "24 8192" "16 16384"; do
set -- $cfg
echo "== BENCH_NBUF=$1 batch $2"
FMIDX_PULL_TRACE=1 BENCH_NBUF=$1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode sharded --big-tokens 400000000 --batch $2 > gpurun_out/r2o_sharded_$1_$2.json 2> gpurun_out/r2o_sharded_$1_$2.err
grep "doc_sharded\|pull trace\|rror" gpurun_out/r2o_sharded_$1_$2.err | tail -3
done
