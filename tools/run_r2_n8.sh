# round 2: the default bench on all 8 GPUs of one box (gpurun --gpus 8): c4 lz4 + gzip, c3, c5 through the product's multi-GPU path
set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2n8_gpus.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2n8_bench.json 2> gpurun_out/r2n8_bench.err
tail -3 gpurun_out/r2n8_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 3 --warmup 3 --no-extra-legs > gpurun_out/r2n4_bench.json 2> gpurun_out/r2n4_bench.err
python - <<'PY'
import json
def load(f):
    txt = open(f).read()
    return json.loads(txt[txt.index('{"metric"'):].strip().splitlines()[0])
for f in ("r2n8_bench", "r2n4_bench"):
    try:
        d = load(f"gpurun_out/{f}.json")
        print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["device_ms_per_step"], d["roofline"]["frac"], d["parity"]["mismatches"], d.get("e2e", {}).get("value"))
        for k in ("gzip", "c3", "c5"):
            if k in d:
                g = d[k]; print("  ", k, g["value"], g["ms_per_step"], g["device_ms_per_step"], g["parity"]["mismatches"], g.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
