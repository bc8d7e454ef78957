set -x
timeout 300 python tools/gpu_time_step.py > gpurun_out/r2l_step.log 2>&1; tail -4 gpurun_out/r2l_step.log
