set -x
nvidia-smi -L
( time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/r2a_tests.log 2>&1
cat gpurun_out/r2a_tests.log
( time timeout 900 python bench.py --gpus 1 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err ) 2>&1 | tail -3
tail -30 gpurun_out/r2a_bench_n1.err
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2a_ref_n1.json 2> gpurun_out/r2a_ref_n1.err ) 2>&1 | tail -3
tail -5 gpurun_out/r2a_ref_n1.err
cat gpurun_out/r2a_ref_n1.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2a_ncu_launch.log 2>&1
tail -3 gpurun_out/r2a_ncu_launch.log
