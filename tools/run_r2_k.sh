# round 2, call K: upload timing, UPGMA tests, quick c4/c3 with pinned result buffers + the faster histogram
set -x
timeout 300 python tools/gpu_time_upload.py > gpurun_out/r2k_upload.log 2>&1; tail -4 gpurun_out/r2k_upload.log
( time timeout 600 python -m pytest tests -m gpu -x -q -k "upgma or tree or fixture or cli_csv" ) > gpurun_out/r2k_tests.log 2>&1
tail -12 gpurun_out/r2k_tests.log
timeout 600 python bench.py --no-gzip-leg --no-extra-legs --no-host-stages --steps 2 --warmup 1 > gpurun_out/r2k_c4.json 2> gpurun_out/r2k_c4.err
timeout 600 python bench.py --config c3 --no-host-stages --steps 2 --warmup 1 > gpurun_out/r2k_c3.json 2> gpurun_out/r2k_c3.err
python - <<'PY'
import json
def load(f):
    txt = open(f).read()
    return json.loads(txt[txt.index('{"metric"'):].strip().splitlines()[0])
for f in ("c4", "c3"):
    try:
        d = load(f"gpurun_out/r2k_{f}.json")
        print(f, d["value"], d["ms_per_step"], d["device_ms_per_step"], d["roofline"]["achieved"], d["parity"]["mismatches"], d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
