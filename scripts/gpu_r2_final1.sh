( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 ) > gpurun_out/r2_final_tests.log 2>&1
cat gpurun_out/r2_final_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
