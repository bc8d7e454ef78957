set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1h_tests.log
python bench.py > gpurun_out/r1h_bench.json 2> gpurun_out/r1h_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1h_launches.csv \
    python bench.py > gpurun_out/r1h_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lz4_pk_pair_kernel -c 1 -f -o gpurun_out/r1h_pk_pair_full \
    python bench.py --genomes 64 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1h_ncu_full.log 2>&1
python bench.py --codec gzip --genomes 64 --steps 1 --warmup 1 > gpurun_out/r1h_gzip64.json 2> gpurun_out/r1h_gzip64.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1h_gzip_launches.csv \
    python bench.py --codec gzip --genomes 64 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1h_ncu_gzip.log 2>&1
tail -c 300 gpurun_out/r1h_tests.log
