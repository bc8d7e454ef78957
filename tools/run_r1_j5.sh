set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2a_tests.log
timeout 600 python bench.py --codec gzip --genomes 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2a_gzip512.json 2> gpurun_out/r2a_gzip512.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "regex:dfl_junction|dfl_parse" -c 60 --csv --log-file gpurun_out/r2a_gzip512_launches.csv \
  python bench.py --codec gzip --genomes 512 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2a_ncu.log 2>&1
cat gpurun_out/r2a_tests.log
