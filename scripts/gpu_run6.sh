set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2f_tests.log
cat gpurun_out/r2f_tests.log
for np in 8 12 16; do BENCH_NPIPE=$np timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('npipe $np value %.1fM e2e %.1fM serial %s abl %.1fM' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['other_kernel_ms'], d['ablation']['unigram_table_off']['value']/1e6))"; done
BENCH_NPIPE=8 BENCH_NSLOT=12 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('nslot12 value %.1fM e2e %.1fM' % (d['value']/1e6, d['e2e']['value']/1e6))"
BENCH_NPIPE=8 BENCH_NSLOT=16 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('nslot16 value %.1fM e2e %.1fM' % (d['value']/1e6, d['e2e']['value']/1e6))"
