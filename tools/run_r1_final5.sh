set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "repetitive" 2>&1 | tail -15 > gpurun_out/r2n_tests.log
cat gpurun_out/r2n_tests.log
