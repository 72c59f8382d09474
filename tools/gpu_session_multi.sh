#!/bin/bash
# Multi-GPU bench arms in one gpurun call (charged N x box time: keep it short).
#
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_session_multi.sh r2m 2'
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_session_multi.sh r2m8 8'
#
# Arms: eager DDP (the default), whole-step CUDA graph with DDP inside, bf16 parameters (bf16 gradient all-reduce); each under
# its own timeout so that a stalled collective costs minutes, not the call (bench.py's watchdog dumps stacks after 7 minutes).
# NCCL_DEBUG=INFO of the first arm goes to <tag>_nccl.log (transport / algorithm selection, NVLS or not).
TAG=${1:-m}
N=${2:-2}
O=gpurun_out
mkdir -p $O
PORT=29511
run() {   # name, extra bench flags...
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 10 --warmup 3 "$@" > $O/${TAG}_bench_${name}.json 2> $O/${TAG}_bench_${name}.err
  echo "$name rc=$?" >> $O/${TAG}_status.txt
  PORT=$((PORT + 1))
}
{ nvidia-smi topo -m; nproc; free -g; df -h /dev/shm; } > $O/${TAG}_env.txt 2>&1
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=$O/${TAG}_nccl.%h.%p.log run eager
run bf16params --bf16-params
run graph --cuda-graph
run graph_bf16params --cuda-graph --bf16-params
SMB_DIR_STREAMS=1 run graph_streams --cuda-graph
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_single.json 2> $O/${TAG}_bench_single.err
cat $O/${TAG}_status.txt
python tools/session_report.py $TAG $O | cut -c1-200
