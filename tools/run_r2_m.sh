# round 2, call M: kernel without the carried skip counter: lz4 parity tests, 104-tile timing, launch list of one c4 step (library kernels only),
# DRAM traffic of the dominant kernels, ncu --set full of the pair kernel
set -x
( time timeout 900 python -m pytest tests -m gpu -x -q -k "lz4 or flagged or non_alphabet or packed or c4 or c3 or fixture or c1" ) > gpurun_out/r2m_tests.log 2>&1; tail -3 gpurun_out/r2m_tests.log
CMD2="python bench.py --genomes 104 --steps 2 --warmup 1 --no-cpu-baseline --no-gzip-leg --no-host-stages --no-e2e"
timeout 300 $CMD2 > gpurun_out/r2m_plain104.json 2> gpurun_out/r2m_plain104.err &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:lz4_pk_pair_kernel -c 1 -f -o gpurun_out/r2m_pk_pair \
    $CMD2 > gpurun_out/r2m_ncu_full.log 2>&1
CMD="python bench.py --steps 1 --warmup 0 --no-extra-legs --no-cpu-baseline --no-host-stages --no-e2e"
timeout 600 $CMD > gpurun_out/r2m_plain_step.json 2> gpurun_out/r2m_plain_step.err &&
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --kernel-name-base demangled \
    -k regex:"lz4_|dfl_|pk_|ncd_|scatter_" --log-file gpurun_out/r2m_launches.csv $CMD > gpurun_out/r2m_ncu_launches.log 2>&1
python - <<'PY'
import json
def load(f):
    txt = open(f).read()
    return json.loads(txt[txt.index('{"metric"'):].strip().splitlines()[0])
for f in ("plain104", "plain_step"):
    d = load(f"gpurun_out/r2m_{f}.json"); print(f, d["value"], d["ms_per_step"], d["device_ms_per_step"], d["roofline"]["achieved"], d["parity"]["mismatches"])
PY
