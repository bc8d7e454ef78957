set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1e_tests.log
python bench.py > gpurun_out/r1e_bench.json 2> gpurun_out/r1e_bench.err
tail -c 600 gpurun_out/r1e_bench.err
