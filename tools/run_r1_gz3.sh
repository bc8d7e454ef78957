set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1n_tests.log
timeout 300 python bench.py --codec gzip --genomes 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1n_gzip64.json 2> gpurun_out/r1n_gzip64.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:dfl_junction3_kernel -c 1 -f -o gpurun_out/r1n_j3 \
    python bench.py --codec gzip --genomes 64 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1n_ncu_j3.log 2>&1
SNACC_DFL_JUNCTION2=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:dfl_junction_kernel -c 1 -f -o gpurun_out/r1n_j2 \
    python bench.py --codec gzip --genomes 64 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1n_ncu_j2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "regex:dfl_" -c 400 --csv --log-file gpurun_out/r1n_gzip_launches.csv \
    python bench.py --codec gzip --genomes 64 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1n_ncu_gzip.log 2>&1
cat gpurun_out/r1n_tests.log
