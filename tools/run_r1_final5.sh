timeout 60 python -m pytest tests -m gpu -x -q -k "fixture" 2>&1 | tail -3 > gpurun_out/r2s_tests.log
cat gpurun_out/r2s_tests.log
