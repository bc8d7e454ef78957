# round 2, call C: ncu --set full of the lean LZ4 pair kernel (104 tiles of 104 streams)
set -x
CMD="python bench.py --genomes 104 --steps 1 --warmup 0 --no-cpu-baseline --no-gzip-leg --no-host-stages --no-e2e"
timeout 300 $CMD > gpurun_out/r2c_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:lz4_pk_pair_kernel -c 1 -f -o gpurun_out/r2c_pk_pair \
    $CMD > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_plain.log | cut -c1-600; tail -5 gpurun_out/r2c_ncu.log
