# round 2: c3 leg alone on all GPUs of the box after the gather fix
set -x
mkdir -p gpurun_out
N=$(nvidia-smi --query-gpu=index --format=csv,noheader | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --config c3 --steps 3 --warmup 3 --no-gzip-leg --no-extra-legs --no-host-stages > gpurun_out/r2n${N}_c3.json 2> gpurun_out/r2n${N}_c3.err; tail -2 gpurun_out/r2n${N}_c3.err
python - <<PY
import json
s=open('gpurun_out/r2n${N}_c3.json').read(); d=json.loads(s[s.index('{"metric"'):].splitlines()[0])
print('c3 n', d['n_gpus'], d['value'], d['ms_per_step'], d['device_ms_per_step'], d['e2e']['value'], d['parity']['mismatches'], d['clocks'])
PY
