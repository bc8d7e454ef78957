set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "lz4 or fixture or regime or small" 2>&1 | tail -4 > gpurun_out/r2d_tests.log
timeout 600 python bench.py --genomes 208 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2d_lz4_208.json 2> gpurun_out/r2d_lz4_208.err
cat gpurun_out/r2d_tests.log
