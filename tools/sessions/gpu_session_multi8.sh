#!/bin/bash
# 8-GPU slot: DDP training step (config 4) and sliding-window inference sharded over the ranks (config 5)
TAG=${1:-r2n}
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/${TAG}_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 10 --warmup 3 > $O/${TAG}_bench_n8.json 2> $O/${TAG}_bench_n8.err; echo "train n8 rc=$?"; cut -c1-300 $O/${TAG}_bench_n8.json; tail -3 $O/${TAG}_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --workload sliding_window --steps 3 --warmup 1 > $O/${TAG}_bench_sw_n8.json 2> $O/${TAG}_bench_sw_n8.err; echo "sw n8 rc=$?"; cut -c1-400 $O/${TAG}_bench_sw_n8.json; tail -3 $O/${TAG}_bench_sw_n8.err
