set -x
K='regex:lz4_|dfl_|pk_|ncd_|scatter_|snacc'
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1m_tests.log
python bench.py --codec gzip --genomes 64 --steps 1 --warmup 1 > gpurun_out/r1m_gzip64.json 2> gpurun_out/r1m_gzip64.err
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "$K" -c 400 --csv --log-file gpurun_out/r1m_gzip_launches.csv \
    python bench.py --codec gzip --genomes 64 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1m_ncu_gzip.log 2>&1
python bench.py --codec gzip --genomes 256 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1m_gzip256.json 2> gpurun_out/r1m_gzip256.err
cat gpurun_out/r1m_tests.log
python bench.py --codec gzip --genomes 512 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1m_gzip512.json 2> gpurun_out/r1m_gzip512.err
