set -x
mkdir -p gpurun_out
timeout 300 python tools/gpu_deflate_prep_only.py 148 > gpurun_out/r2e_prep.log 2>&1; cat gpurun_out/r2e_prep.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dfl_|radix" -c 200 --csv --log-file gpurun_out/r2e_launches.csv python tools/gpu_deflate_prep_only.py 148 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dfl_prep_kernel -c 1 -o gpurun_out/r2e_prep_ncu -f python tools/gpu_deflate_prep_only.py 148 > gpurun_out/r2e_ncu.log 2>&1; tail -3 gpurun_out/r2e_ncu.log
