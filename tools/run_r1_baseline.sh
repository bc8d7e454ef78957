set -x
python -m pytest tests -m gpu -x -q -k "lz4 and not gzip" 2>&1 | tail -5 > gpurun_out/r1_tests.log
python bench.py --steps 2 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
SMALL="python bench.py --genomes 64 --length 1000000 --rows 4 --steps 1 --warmup 1 --no-cpu-baseline"
$SMALL > gpurun_out/r1_small.json 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches.csv $SMALL > gpurun_out/r1_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lz4_stream -s 2 -c 1 -o gpurun_out/r1_lz4_stream $SMALL > gpurun_out/r1_ncu2.log 2>&1
ls -la gpurun_out
