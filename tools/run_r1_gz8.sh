set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1v_tests.log
timeout 600 python bench.py --codec gzip --genomes 64 --steps 1 --warmup 1 > gpurun_out/r1v_gzip64.json 2> gpurun_out/r1v_gzip64.err
timeout 600 python bench.py --codec gzip --genomes 512 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1v_gzip512.json 2> gpurun_out/r1v_gzip512.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "regex:dfl_" -c 1000 --csv --log-file gpurun_out/r1v_gzip1024_launches.csv \
  python bench.py --config c5 --genomes 1024 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1v_ncu_gzip1024.log 2>&1
timeout 900 python bench.py --config c5 --genomes 1024 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1v_gzip1024.json 2> gpurun_out/r1v_gzip1024.err
cat gpurun_out/r1v_tests.log
