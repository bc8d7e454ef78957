# round 2, final multi-GPU run at HEAD: default bench on all GPUs of the box (product path, NCCL)
set -x
mkdir -p gpurun_out
N=$(nvidia-smi --query-gpu=index --format=csv,noheader | wc -l)
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2n${N}_gpus.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2n${N}_bench.json 2> gpurun_out/r2n${N}_bench.err; tail -3 gpurun_out/r2n${N}_bench.err
python - <<PY
import json
s=open('gpurun_out/r2n${N}_bench.json').read(); d=json.loads(s[s.index('{"metric"'):].splitlines()[0])
print('n', d['n_gpus'], d['value'], d['ms_per_step'], d['device_ms_per_step'], d['e2e']['value'], d['parity']['mismatches'])
for k in ('gzip','c3','c5'): print('  ', k, d[k]['value'], d[k]['ms_per_step'], d[k]['device_ms_per_step'], d[k]['e2e']['value'], d[k]['parity']['mismatches'])
PY
