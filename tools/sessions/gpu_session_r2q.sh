#!/bin/bash
TAG=${1:-r2q}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_instnorm.py tests/test_gpu_fullsize.py tests/test_gpu_model.py -q -x -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log | cut -c1-200
for f in 1 0; do
  SMB_IN_FUSED=$f timeout 200 python tools/op_breakdown.py > $O/${TAG}_breakdown_fused$f.txt 2>&1
  echo "SMB_IN_FUSED=$f small-shape instnorm ms/step:"; grep instnorm $O/${TAG}_breakdown_fused$f.txt | awk '{ if ($7+0 <= 262144) s+=$1; else b+=$1 } END {print "  small:", s, "big:", b}'
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-230 $O/${TAG}_bench.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda --cuda-graph > $O/${TAG}_bench_graph.json 2> $O/${TAG}_bench_graph.err; echo "graph rc=$?"; cut -c1-230 $O/${TAG}_bench_graph.json; tail -2 $O/${TAG}_bench_graph.err
