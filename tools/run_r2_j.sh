# round 2, call J (stuck-lane general steps, EXC unroll 4): flagged-base kernels (dirty ring inside the ring, 104 streams): parity tests, clean vs flagged timing, c4 with 1e-4 flagged
set -x
( time timeout 1200 python -m pytest tests -m gpu -x -q -k "lz4 or flagged or non_alphabet or packed or c4 or c3 or fixture" ) > gpurun_out/r2j_tests.log 2>&1
tail -15 gpurun_out/r2j_tests.log
CMD="python bench.py --genomes 104 --steps 2 --warmup 1 --no-cpu-baseline --no-gzip-leg --no-host-stages --no-e2e"
timeout 300 $CMD > gpurun_out/r2j_clean.json 2> gpurun_out/r2j_clean.err
timeout 300 $CMD --exceptions 1e-5 > gpurun_out/r2j_exc5.json 2> gpurun_out/r2j_exc5.err
timeout 300 $CMD --exceptions 1e-4 > gpurun_out/r2j_exc4.json 2> gpurun_out/r2j_exc4.err
timeout 300 $CMD --exceptions 1e-3 > gpurun_out/r2j_exc3.json 2> gpurun_out/r2j_exc3.err
timeout 600 python bench.py --exceptions 1e-4 --no-gzip-leg --no-extra-legs --no-host-stages --steps 2 --warmup 1 > gpurun_out/r2j_c4exc4.json 2> gpurun_out/r2j_c4exc4.err
python - <<'PY'
import json
def load(f):
    txt = open(f).read()
    return json.loads(txt[txt.index('{"metric"'):].strip().splitlines()[0])
for f in ("clean", "exc5", "exc4", "exc3", "c4exc4"):
    try:
        d = load(f"gpurun_out/r2j_{f}.json")
        print(f, d["value"], d["ms_per_step"], d["device_ms_per_step"], d["roofline"]["achieved"], d["parity"]["mismatches"], d["packed_jobs_per_step"], d["bytewise_jobs_per_step"], d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
