set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2m_tests.log
timeout 900 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
timeout 300 python bench.py --codec gzip --genomes 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2m_gzip512_base.json 2> gpurun_out/r2m_gzip512_base.err
SNACC_B200_LIB=build/libsnacc_cumdirect.so timeout 300 python bench.py --codec gzip --genomes 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2m_gzip512_cumdirect.json 2> gpurun_out/r2m_gzip512_cumdirect.err
cat gpurun_out/r2m_tests.log
