set -x
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:dfl_parse_kernel -c 1 -f -o gpurun_out/r1s_parse \
    python bench.py --codec gzip --genomes 256 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1s_ncu_parse.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "regex:dfl_parse|dfl_junction|dfl_prep|dfl_match" -c 100 --csv --log-file gpurun_out/r1s_gzip512_launches.csv \
    python bench.py --codec gzip --genomes 512 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1s_ncu_gzip512.log 2>&1
