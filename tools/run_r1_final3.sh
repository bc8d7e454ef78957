set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2g_tests.log
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "regex:lz4_|dfl_|pk_|ncd_|scatter_" -c 400 --csv --log-file gpurun_out/r2g_launches.csv \
    python bench.py > gpurun_out/r2g_ncu_bench.log 2>&1
cat gpurun_out/r2g_tests.log; tail -3 gpurun_out/r2g_smoke.log
