#!/bin/bash
TAG=${1:-r2k}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/${TAG}_pytest.log | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-260 $O/${TAG}_bench.json; tail -2 $O/${TAG}_bench.err
SMB_GEMM=all timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda > $O/${TAG}_bench_gemmall.json 2> $O/${TAG}_bench_gemmall.err; cut -c1-260 $O/${TAG}_bench_gemmall.json
SMB_GEMM=off timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda > $O/${TAG}_bench_gemmoff.json 2> $O/${TAG}_bench_gemmoff.err; cut -c1-260 $O/${TAG}_bench_gemmoff.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda --cuda-graph > $O/${TAG}_bench_graph.json 2> $O/${TAG}_bench_graph.err; cut -c1-260 $O/${TAG}_bench_graph.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda --amp fp16 > $O/${TAG}_bench_fp16.json 2> $O/${TAG}_bench_fp16.err; cut -c1-260 $O/${TAG}_bench_fp16.json
