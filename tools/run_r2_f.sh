# round 2, call F: flagged-base (EXC) kernels: parity tests, timing clean vs 1e-4 flagged at 104 genomes, default bench incl. c3/c5 legs
set -x
( time timeout 1200 python -m pytest tests -m gpu -x -q -k "lz4 or flagged or non_alphabet or packed or c4 or c3 or fixture" ) > gpurun_out/r2f_tests.log 2>&1
tail -15 gpurun_out/r2f_tests.log
CMD="python bench.py --genomes 104 --steps 2 --warmup 1 --no-cpu-baseline --no-gzip-leg --no-host-stages --no-e2e"
timeout 300 $CMD > gpurun_out/r2f_clean.json 2> gpurun_out/r2f_clean.err
timeout 300 $CMD --exceptions 1e-4 > gpurun_out/r2f_exc4.json 2> gpurun_out/r2f_exc4.err
timeout 300 $CMD --exceptions 1e-5 > gpurun_out/r2f_exc5.json 2> gpurun_out/r2f_exc5.err
timeout 1200 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -3 gpurun_out/r2f_bench.err
python - <<'PY'
import json
def load(f):
    txt = open(f).read()
    return json.loads(txt[txt.index('{"metric"'):].strip().splitlines()[0])
for f in ("clean", "exc4", "exc5", "bench"):
    try:
        d = load(f"gpurun_out/r2f_{f}.json")
        print(f, d["value"], d["ms_per_step"], d["device_ms_per_step"], d["roofline"]["achieved"], d["parity"], d["packed_jobs_per_step"], d["bytewise_jobs_per_step"], d.get("e2e", {}).get("value"))
        for k in ("gzip", "c3", "c5"):
            if k in d:
                g = d[k]; print("  ", k, g["value"], g["ms_per_step"], g["parity"], g.get("e2e", {}).get("value"))
        if "host_s" in d: print("  host_s", d["host_s"])
    except Exception as e:
        print(f, "failed", e)
PY
