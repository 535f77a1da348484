nvidia-smi -L | wc -l
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 > gpurun_out/r2u_bench_n8.json 2> gpurun_out/r2u_bench_n8.err ) 2>&1 | tail -3
grep -v "NCCL\|^$\|OMP_NUM\|\*\*\*\*" gpurun_out/r2u_bench_n8.err | grep "rank 0\|rror\|trace" | tail -12
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2u_bench_n8.json").read().strip().splitlines()[-1])
print("value %.1fM ms/step %.4f e2e %.1fM launches %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["gpu_launches"]))
for k,v in d["modes"].items(): print(k, {a:b for a,b in v.items() if a not in ("what","parity_note","exchange")} if isinstance(v, dict) else v)
PY
