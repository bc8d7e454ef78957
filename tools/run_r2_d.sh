# round 2, call D: lean loop v3 (branch-free, unrolled, full-warp votes): parity tests, 104-tile timing (4x26 and 8x13 lanes),
# singles pass lean vs speculative, quick c4, ncu --set full
set -x
( time timeout 900 python -m pytest tests -m gpu -x -q -k "lz4 or c1 or c3 or c4 or fixture or stale or packed or repetitive or regime or small_stream" ) > gpurun_out/r2d_tests.log 2>&1
tail -3 gpurun_out/r2d_tests.log
CMD="python bench.py --genomes 104 --steps 2 --warmup 1 --no-cpu-baseline --no-gzip-leg --no-host-stages --no-e2e"
timeout 300 $CMD --opt lz4_lanes=13 > gpurun_out/r2d_l13.json 2> gpurun_out/r2d_l13.err
timeout 300 $CMD --opt lz4_singles_lean=1 > gpurun_out/r2d_sl.json 2> gpurun_out/r2d_sl.err
timeout 300 $CMD > gpurun_out/r2d_plain.json 2> gpurun_out/r2d_plain.err &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:lz4_pk_pair_kernel -c 1 -f -o gpurun_out/r2d_pk_pair \
    $CMD > gpurun_out/r2d_ncu.log 2>&1
timeout 600 python bench.py --no-gzip-leg --no-host-stages --steps 2 --warmup 1 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python - <<'PY'
import json
for f in ("l13", "sl", "plain", "bench"):
    try:
        d = json.load(open(f"gpurun_out/r2d_{f}.json"))
        print(f, d["value"], d["ms_per_step"], d["device_ms_per_step"], d["roofline"]["achieved"], d["parity"]["mismatches"], d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2d_ncu.log
