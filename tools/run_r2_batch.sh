set -x
mkdir -p gpurun_out
timeout 300 python tools/gpu_singles_only.py 148 > gpurun_out/r2b_singles.log 2>&1; cat gpurun_out/r2b_singles.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lz4 or packed or flagged or c4 or c1 or c2 or segments or regime or repetitive or stale or c3 or fixture" > gpurun_out/r2b_tests.log 2>&1; tail -3 gpurun_out/r2b_tests.log
