set -x
K='regex:lz4_|dfl_|pk_|ncd_|scatter_|snacc'
python bench.py --genomes 104 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1i_l26.json 2> gpurun_out/r1i_l26.err
SNACC_PK_LANES=13 python bench.py --genomes 104 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1i_l13.json 2> gpurun_out/r1i_l13.err
SNACC_PK_LANES=13 python bench.py --genomes 208 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1i_l13_208.json 2> gpurun_out/r1i_l13_208.err
python bench.py --genomes 208 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1i_l26_208.json 2> gpurun_out/r1i_l26_208.err
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "$K" -c 400 --csv --log-file gpurun_out/r1i_gzip_launches.csv \
    python bench.py --codec gzip --genomes 64 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1i_ncu_gzip.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "$K" -c 400 --csv --log-file gpurun_out/r1i_launches.csv \
    python bench.py > gpurun_out/r1i_ncu_bench.log 2>&1
