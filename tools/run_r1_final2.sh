set -x
timeout 600 python bench.py --codec zlib --genomes 64 --steps 1 --warmup 1 > gpurun_out/r2c_zlib64.json 2> gpurun_out/r2c_zlib64.err
timeout 600 python bench.py --codec zlib --genomes 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2c_zlib256.json 2> gpurun_out/r2c_zlib256.err
timeout 1200 python bench.py --config c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2c_c5.json 2> gpurun_out/r2c_c5.err
tail -c 300 gpurun_out/r2c_c5.err
