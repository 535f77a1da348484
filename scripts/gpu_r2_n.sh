for cfg in "8 0" "16 0" "24 0" "24 32"; do
set -- $cfg
echo "== BENCH_NBUF=$1 FMIDX_PULL_CTAS=$2"
FMIDX_PULL_TRACE=1 BENCH_NBUF=$1 FMIDX_PULL_CTAS=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode sharded --big-tokens 400000000 > gpurun_out/r2n_sharded_$1_$2.json 2> gpurun_out/r2n_sharded_$1_$2.err
grep "doc_sharded\|pull trace\|rror" gpurun_out/r2n_sharded_$1_$2.err | tail -3
done
