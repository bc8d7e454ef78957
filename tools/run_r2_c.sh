# round 2, call C: lean LZ4 pair kernel, epoch bits by plain stores (default lib) vs shared atomics (build/exp/lib_atomic.so);
# then ncu --set full of the default (104 tiles of 104 streams)
set -x
CMD="python bench.py --genomes 104 --steps 2 --warmup 1 --no-cpu-baseline --no-gzip-leg --no-host-stages --no-e2e"
SNACC_B200_LIB=$PWD/build/exp/lib_atomic.so timeout 300 $CMD > gpurun_out/r2c_atomic.json 2> gpurun_out/r2c_atomic.err
timeout 300 $CMD > gpurun_out/r2c_plain.json 2> gpurun_out/r2c_plain.err &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:lz4_pk_pair_kernel -c 1 -f -o gpurun_out/r2c_pk_pair \
    $CMD > gpurun_out/r2c_ncu.log 2>&1
python - <<'PY'
import json
for f in ("atomic", "plain"):
    try:
        d = json.load(open(f"gpurun_out/r2c_{f}.json"))
        print(f, d["value"], d["device_ms_per_step"], d["roofline"]["achieved"], d["parity"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2c_ncu.log
