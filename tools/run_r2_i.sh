# round 2, call I: ncu --set full of the EXC pair kernel (104 genomes, 1e-5 flagged bases)
set -x
CMD="python bench.py --genomes 104 --steps 1 --warmup 0 --no-cpu-baseline --no-gzip-leg --no-host-stages --no-e2e --exceptions 1e-5"
timeout 300 $CMD > gpurun_out/r2i_plain.json 2> gpurun_out/r2i_plain.err &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:lz4_pk_pair_kernel -c 1 -f -o gpurun_out/r2i_pk_pair_exc \
    $CMD > gpurun_out/r2i_ncu.log 2>&1
tail -3 gpurun_out/r2i_ncu.log
