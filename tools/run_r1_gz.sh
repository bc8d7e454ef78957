set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1c_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1c_smoke.log 2>&1
G=8 L=4500000 python tools/gpu_check_deflate.py > gpurun_out/r1c_dfl.log 2>&1
