#!/bin/bash
# r2b: first hardware run of the tcgen05 GEMM (bounded by timeouts), full-size parity tests, bench with the flipped defaults
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -x -p no:cacheprovider > $O/${TAG}_pytest_gemm.log 2>&1; echo "gemm pytest rc=$?"
tail -25 $O/${TAG}_pytest_gemm.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_train_extras.py -q -s -p no:cacheprovider > $O/${TAG}_pytest_full.log 2>&1; echo "fullsize pytest rc=$?"
tail -15 $O/${TAG}_pytest_full.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda --cuda-graph > $O/${TAG}_bench_graph.json 2> $O/${TAG}_bench_graph.err
timeout 600 python bench.py --workload sliding_window --steps 2 --warmup 1 > $O/${TAG}_bench_sw.json 2> $O/${TAG}_bench_sw.err; echo "sw rc=$?"
for f in $O/${TAG}_bench*.json; do echo "== $f"; cut -c1-600 $f; done
