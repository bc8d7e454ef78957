set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1f_tests.log
python bench.py --config c3 --steps 1 --warmup 1 > gpurun_out/r1f_bench_c3.json 2> gpurun_out/r1f_bench_c3.err
tail -c 800 gpurun_out/r1f_bench_c3.err
