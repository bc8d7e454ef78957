set -x
python -m pytest tests -m gpu -x -q -k "lz4 and not gzip" 2>&1 | tail -5 > gpurun_out/r1b_tests.log
python bench.py > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err
tail -c 600 gpurun_out/r1b_bench.err
SMALL="python bench.py --genomes 96 --length 1000000 --steps 1 --warmup 1 --no-cpu-baseline"
$SMALL > gpurun_out/r1b_small.json 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"snacc|ncd_kernel|scatter_records" -c 60 --csv --log-file gpurun_out/r1b_launches.csv $SMALL > gpurun_out/r1b_ncu1.log 2>&1
python bench.py --impl reference > gpurun_out/r1b_ref.json 2>&1
