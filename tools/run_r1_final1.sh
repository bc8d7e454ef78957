set -x
timeout 900 python bench.py > gpurun_out/r1x_bench.json 2> gpurun_out/r1x_bench.err
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:lz4_pk_pair_kernel -c 1 --csv --log-file gpurun_out/r1x_traffic_lz4.csv \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1x_ncu_traffic.log 2>&1
timeout 600 python bench.py --impl reference > gpurun_out/r1x_ref.json 2> gpurun_out/r1x_ref.err
timeout 900 python bench.py --config c3 --steps 1 --warmup 1 > gpurun_out/r1x_c3.json 2> gpurun_out/r1x_c3.err
tail -c 300 gpurun_out/r1x_bench.err
