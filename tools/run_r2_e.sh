# round 2, call E (gpurun --gpus 2): whole GPU suite incl. the 2-rank NCCL tests, unroll 4 vs 8, 2-GPU bench through the product path
set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2e_gpus.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2e_tests.log 2>&1
tail -4 gpurun_out/r2e_tests.log
CMD="python bench.py --genomes 104 --steps 2 --warmup 1 --no-cpu-baseline --no-gzip-leg --no-host-stages --no-e2e"
SNACC_B200_LIB=$PWD/build/exp/lib_u8.so timeout 300 $CMD > gpurun_out/r2e_u8.json 2> gpurun_out/r2e_u8.err
timeout 300 $CMD > gpurun_out/r2e_u4.json 2> gpurun_out/r2e_u4.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
tail -3 gpurun_out/r2e_bench_n2.err
python - <<'PY'
import json
for f in ("u8", "u4", "bench_n2"):
    try:
        d = json.load(open(f"gpurun_out/r2e_{f}.json"))
        print(f, d["value"], d["ms_per_step"], d["device_ms_per_step"], d["roofline"]["achieved"], d["parity"], d.get("e2e", {}).get("value"))
        if "gzip" in d:
            g = d["gzip"]; print("  gzip", g["value"], g["ms_per_step"], g["parity"], g.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
