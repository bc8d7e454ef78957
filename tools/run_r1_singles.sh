set -x
timeout 300 python tools/gpu_time_singles.py 148 > gpurun_out/r2h_singles_turbo.log 2>&1
SNACC_B200_LIB=build/libsnacc_general.so timeout 300 python tools/gpu_time_singles.py 148 > gpurun_out/r2h_singles_general.log 2>&1
cat gpurun_out/r2h_singles_turbo.log gpurun_out/r2h_singles_general.log
