# round 2, call B: lean LZ4 pair loop -- lz4 parity tests + quick bench
set -x
( time timeout 900 python -m pytest tests -m gpu -x -q -k "lz4 or c1 or c3 or c4 or fixture or stale or packed or repetitive or regime or small_stream" ) > gpurun_out/r2b_tests.log 2>&1
tail -5 gpurun_out/r2b_tests.log
timeout 600 python bench.py --no-gzip-leg --no-host-stages --steps 2 --warmup 1 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 1500 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
timeout 300 python bench.py --config c3 --no-host-stages --steps 1 --warmup 1 > gpurun_out/r2b_c3.json 2> gpurun_out/r2b_c3.err; tail -c 600 gpurun_out/r2b_c3.json; tail -5 gpurun_out/r2b_c3.err
