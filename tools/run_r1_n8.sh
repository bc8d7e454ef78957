set -x
# c4 lz4 at 8 GPUs:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2k_c5_n8.json 2> gpurun_out/r2k_c5_n8.err
tail -c 600 gpurun_out/r2k_c5_n8.err
