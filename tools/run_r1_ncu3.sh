set -x
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:lz4_pk_pair_kernel -c 1 -f -o gpurun_out/r2r_pk_pair_final \
    python bench.py --genomes 104 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2r_ncu.log 2>&1
