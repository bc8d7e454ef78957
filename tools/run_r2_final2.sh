# round 2, final 1-GPU evidence run at HEAD: whole GPU suite, smoke, default bench + reference arm, launch list (ncu) of one c4 step,
# ncu --set full of the LZ4 pair kernel (104 full tiles) and of the singles kernel
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r2z_gpu.txt; nproc >> gpurun_out/r2z_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2z_tests.log 2>&1; tail -4 gpurun_out/r2z_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -2 gpurun_out/r2z_smoke.log
timeout 1500 python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; tail -2 gpurun_out/r2z_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_reference_n1.json 2> gpurun_out/r2z_reference_n1.err
CMD="python bench.py --steps 1 --warmup 0 --no-extra-legs --no-cpu-baseline --no-host-stages --no-e2e"
timeout 600 $CMD > gpurun_out/r2z_plain_step.json 2> gpurun_out/r2z_plain_step.err &&
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"lz4_|dfl_|pk_|ncd_|scatter_" -c 700 --csv --kernel-name-base demangled \
    --log-file gpurun_out/r2z_launches.csv $CMD > gpurun_out/r2z_ncu_launches.log 2>&1
CMD2="python bench.py --genomes 104 --steps 1 --warmup 0 --no-cpu-baseline --no-gzip-leg --no-extra-legs --no-host-stages --no-e2e"
timeout 300 $CMD2 > gpurun_out/r2z_plain104.json 2> gpurun_out/r2z_plain104.err &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:lz4_pk_pair_kernel -c 1 -f -o gpurun_out/r2z_pk_pair \
    $CMD2 > gpurun_out/r2z_ncu_full.log 2>&1
tail -2 gpurun_out/r2z_ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:lz4_pk_single -c 1 -f -o gpurun_out/r2z_pk_single \
    python tools/gpu_singles_only.py 148 > gpurun_out/r2z_ncu_single.log 2>&1
tail -2 gpurun_out/r2z_ncu_single.log
python - <<'PY'
import json
s=open('gpurun_out/r2z_bench_n1.json').read(); d=json.loads(s[s.index('{"metric"'):].splitlines()[0])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['parity'])
for k in ('gzip','c3','c5'): print(k, d[k]['value'], d[k]['ms_per_step'], d[k]['e2e']['value'], d[k].get('parity'))
PY
