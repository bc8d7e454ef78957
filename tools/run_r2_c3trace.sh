set -x
mkdir -p gpurun_out
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --config c3 --steps 3 --warmup 1 --no-cpu-baseline --no-gzip-leg --no-extra-legs --no-host-stages --no-e2e"
SNACC_BENCH_TRACE=1 SNACC_BENCH_NO_SAMPLER=1 timeout 600 $B > gpurun_out/r2t_a.json 2> gpurun_out/r2t_a.err; echo "--- no sampler"; grep "trace rank 0" gpurun_out/r2t_a.err
SNACC_BENCH_TRACE=1 SNACC_BENCH_STEP_BARRIER=1 timeout 600 $B > gpurun_out/r2t_b.json 2> gpurun_out/r2t_b.err; echo "--- barrier per step"; grep "trace rank 0" gpurun_out/r2t_b.err
SNACC_BENCH_TRACE=1 SNACC_BENCH_DROP_REFS=1 timeout 600 $B > gpurun_out/r2t_c.json 2> gpurun_out/r2t_c.err; echo "--- drop refs"; grep "trace rank 0" gpurun_out/r2t_c.err
