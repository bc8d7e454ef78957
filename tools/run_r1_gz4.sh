set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1r_tests.log
timeout 300 python bench.py --codec gzip --genomes 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1r_gzip64.json 2> gpurun_out/r1r_gzip64.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "regex:dfl_" -c 400 --csv --log-file gpurun_out/r1r_gzip_launches.csv \
    python bench.py --codec gzip --genomes 64 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1r_ncu_gzip.log 2>&1
timeout 600 python bench.py --codec gzip --genomes 256 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1r_gzip256.json 2> gpurun_out/r1r_gzip256.err
timeout 600 python bench.py --codec gzip --genomes 512 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1r_gzip512.json 2> gpurun_out/r1r_gzip512.err
cat gpurun_out/r1r_tests.log
