set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "fast_mode or cli_csv or fixture" 2>&1 | tail -15 > gpurun_out/r2o_tests.log
cat gpurun_out/r2o_tests.log
