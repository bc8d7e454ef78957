set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gzip or zlib or deflate or c5 or c2 or fixture or regime or repetitive or small_stream" > gpurun_out/r2f_tests.log 2>&1; tail -5 gpurun_out/r2f_tests.log
timeout 600 python tools/gpu_time_deflate_prep.py > gpurun_out/r2f_prep.log 2>&1; cat gpurun_out/r2f_prep.log
