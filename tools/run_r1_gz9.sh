set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1w_tests.log
timeout 600 python bench.py --codec gzip --genomes 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1w_gzip512.json 2> gpurun_out/r1w_gzip512.err
timeout 1200 python bench.py --config c5 --steps 1 --warmup 1 > gpurun_out/r1w_c5.json 2> gpurun_out/r1w_c5.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv > gpurun_out/r1w_mem.log
cat gpurun_out/r1w_tests.log; tail -c 400 gpurun_out/r1w_c5.err
