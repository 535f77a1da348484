timeout 300 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_parity.py -m gpu -x -q -k "graph or host or pipelin or summary or submit or golden or sample" 2>&1 | tail -4
for cfg in "8 0" "16 0" "16 296"; do
set -- $cfg
echo "== BENCH_NBUF=$1 FMIDX_PULL_CTAS=$2"
FMIDX_PULL_TRACE=1 BENCH_NBUF=$1 FMIDX_PULL_CTAS=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode sharded --big-tokens 400000000 > gpurun_out/r2m_sharded_$1_$2.json 2> gpurun_out/r2m_sharded_$1_$2.err
grep "doc_sharded\|pull trace" gpurun_out/r2m_sharded_$1_$2.err | tail -3
done
