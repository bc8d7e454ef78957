# round 2: 2-GPU check at HEAD: the two-rank NCCL tests, then the default bench on 2 ranks
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2n2_gpus.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_rank" > gpurun_out/r2n2_tests.log 2>&1; tail -3 gpurun_out/r2n2_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2n2_bench.json 2> gpurun_out/r2n2_bench.err; tail -3 gpurun_out/r2n2_bench.err
python - <<'PY'
import json
s=open('gpurun_out/r2n2_bench.json').read(); d=json.loads(s[s.index('{"metric"'):].splitlines()[0])
print('r2n2', d['n_gpus'], d['value'], d['ms_per_step'], d['device_ms_per_step'], d['e2e']['value'], d['parity'])
for k in ('gzip','c3','c5'): print('  ', k, d[k]['value'], d[k]['ms_per_step'], d[k]['device_ms_per_step'], d[k]['e2e']['value'], d[k].get('parity'))
PY
