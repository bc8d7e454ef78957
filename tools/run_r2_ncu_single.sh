set -x
mkdir -p gpurun_out
timeout 300 python tools/gpu_singles_only.py 148 > gpurun_out/r2c_singles.log 2>&1; cat gpurun_out/r2c_singles.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lz4_pk_single -c 1 -o gpurun_out/r2c_single_ncu -f python tools/gpu_singles_only.py 148 > gpurun_out/r2c_ncu.log 2>&1; tail -3 gpurun_out/r2c_ncu.log
ls -la gpurun_out/*.ncu-rep
