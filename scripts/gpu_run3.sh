set -x
python benchmarks/prof_step.py --steps 8 > gpurun_out/r2c_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_search|k_mask' -s 12 -c 32 --csv --log-file gpurun_out/r2c_launches.csv python benchmarks/prof_step.py --steps 8 > gpurun_out/r2c_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_search|k_mask' -s 12 -c 8 -o gpurun_out/r2c_prof python benchmarks/prof_step.py --steps 8 > gpurun_out/r2c_ncu2.log 2>&1
ls -la gpurun_out/r2c*
