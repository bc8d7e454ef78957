set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lz4 or packed or flagged or c4 or c1 or c2 or segments" > gpurun_out/r2s_tests.log 2>&1; tail -3 gpurun_out/r2s_tests.log
timeout 600 python tools/gpu_time_segments.py > gpurun_out/r2s_seg.log 2>&1; cat gpurun_out/r2s_seg.log
timeout 600 python tools/gpu_time_segments.py 0.0001 > gpurun_out/r2s_seg_exc.log 2>&1; cat gpurun_out/r2s_seg_exc.log
