set -x
timeout 600 python bench.py --codec gzip --genomes 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1y_ovl.json 2> gpurun_out/r1y_ovl.err
SNACC_DFL_NO_OVERLAP=1 timeout 600 python bench.py --codec gzip --genomes 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1y_noovl.json 2> gpurun_out/r1y_noovl.err
