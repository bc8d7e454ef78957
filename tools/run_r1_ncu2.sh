set -x
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:dfl_junction_kernel -c 1 -f -o gpurun_out/r1z_junction \
    python bench.py --codec gzip --genomes 128 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1z_ncu_j.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:dfl_parse_kernel -c 1 -f -o gpurun_out/r1z_parse \
    python bench.py --codec gzip --genomes 256 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1z_ncu_p.log 2>&1
