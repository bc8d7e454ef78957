set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2e_lz4_n2.json 2> gpurun_out/r2e_lz4_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c5 --genomes 1024 --steps 1 --warmup 1 > gpurun_out/r2e_gzip1024_n2.json 2> gpurun_out/r2e_gzip1024_n2.err
tail -c 400 gpurun_out/r2e_lz4_n2.err; tail -c 400 gpurun_out/r2e_gzip1024_n2.err
