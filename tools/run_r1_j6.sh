set -x
timeout 600 python bench.py --codec gzip --genomes 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_gzip512.json 2> gpurun_out/r2b_gzip512.err
