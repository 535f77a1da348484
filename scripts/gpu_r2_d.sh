nvidia-smi -L | wc -l
( timeout 600 python -m pytest tests/test_multigpu.py tests/test_gpu_robustness.py -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2d_mgpu_tests.log 2>&1
cat gpurun_out/r2d_mgpu_tests.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err ) 2>&1 | tail -3
grep -v "NCCL\|^$\|OMP_NUM\|\*\*\*\*" gpurun_out/r2d_bench_n2.err | tail -15
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2d_bench_n2.json").read().strip().splitlines()[-1])
print("value %.1fM ms/step %.4f e2e %.1fM launches %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["gpu_launches"]))
for k,v in d["modes"].items(): print(k, {a:b for a,b in v.items() if a not in ("what","parity_note","exchange")})
PY
