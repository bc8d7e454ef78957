set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2q_tests.log
timeout 900 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
cat gpurun_out/r2q_tests.log
