# round 2, mid-round check at HEAD: whole GPU suite, smoke, default bench
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2y_tests.log 2>&1; tail -4 gpurun_out/r2y_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2y_smoke.log 2>&1; tail -2 gpurun_out/r2y_smoke.log
timeout 1500 python bench.py > gpurun_out/r2y_bench_n1.json 2> gpurun_out/r2y_bench_n1.err; tail -2 gpurun_out/r2y_bench_n1.err
python - <<'PY'
import json
s=open('gpurun_out/r2y_bench_n1.json').read(); d=json.loads(s[s.index('{"metric"'):].splitlines()[0])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['parity'])
for k in ('gzip','c3','c5'): print(k, d[k]['value'], d[k]['ms_per_step'], d[k].get('parity'))
PY
