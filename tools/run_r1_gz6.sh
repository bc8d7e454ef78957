set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1t_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "regex:dfl_" -c 400 --csv --log-file gpurun_out/r1t_gzip256_launches.csv \
    python bench.py --codec gzip --genomes 256 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1t_ncu_gzip256.log 2>&1
timeout 600 python bench.py --codec gzip --genomes 512 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1t_gzip512.json 2> gpurun_out/r1t_gzip512.err
timeout 900 python bench.py --config c5 --genomes 1024 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1t_gzip1024.json 2> gpurun_out/r1t_gzip1024.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv > gpurun_out/r1t_mem.log
cat gpurun_out/r1t_tests.log
