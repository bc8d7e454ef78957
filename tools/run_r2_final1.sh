# round 2, final 1-GPU evidence run: whole GPU suite, smoke, default bench + reference arm, launch list (ncu) of the bench command,
# ncu --set full of the LZ4 pair kernel, DRAM traffic of the dominant kernels
set -x
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r2z_gpu.txt; nproc >> gpurun_out/r2z_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2z_tests.log 2>&1; tail -4 gpurun_out/r2z_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -2 gpurun_out/r2z_smoke.log
timeout 1500 python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; tail -2 gpurun_out/r2z_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_reference_n1.json 2> gpurun_out/r2z_reference_n1.err
# launch list of one c4 step (lz4 + gzip legs): the same command first without ncu
CMD="python bench.py --steps 1 --warmup 0 --no-extra-legs --no-cpu-baseline --no-host-stages --no-e2e"
timeout 600 $CMD > gpurun_out/r2z_plain_step.json 2> gpurun_out/r2z_plain_step.err &&
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --kernel-name-base demangled \
    --log-file gpurun_out/r2z_launches.csv $CMD > gpurun_out/r2z_ncu_launches.log 2>&1
# the dominant kernel, full set, on full tiles (104 genomes: 104 tiles of 104 streams)
CMD2="python bench.py --genomes 104 --steps 1 --warmup 0 --no-cpu-baseline --no-gzip-leg --no-host-stages --no-e2e"
timeout 300 $CMD2 > gpurun_out/r2z_plain104.json 2> gpurun_out/r2z_plain104.err &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:lz4_pk_pair_kernel -c 1 -f -o gpurun_out/r2z_pk_pair \
    $CMD2 > gpurun_out/r2z_ncu_full.log 2>&1
tail -2 gpurun_out/r2z_ncu_full.log
