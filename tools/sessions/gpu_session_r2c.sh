#!/bin/bash
TAG=${1:-r2c}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -p no:cacheprovider > $O/${TAG}_pytest_gemm.log 2>&1; echo "gemm pytest rc=$?"; tail -12 $O/${TAG}_pytest_gemm.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_scan.py tests/test_gpu_scan_variants.py tests/test_gpu_fullsize.py tests/test_gpu_model.py -q -p no:cacheprovider > $O/${TAG}_pytest_scan.log 2>&1; echo "scan pytest rc=$?"; tail -12 $O/${TAG}_pytest_scan.log | cut -c1-200
timeout 300 python tools/microbench.py --dtypes bf16,f32 --batches 2 --no-ref --out $O/${TAG}_mb_dense.json > $O/${TAG}_mb_dense.log 2>&1
timeout 300 python tools/microbench.py --dtypes bf16 --batches 2 --no-ref --no-dense --out $O/${TAG}_mb_nodense.json > $O/${TAG}_mb_nodense.log 2>&1
grep -h scan_bwd_ms $O/${TAG}_mb_dense.log $O/${TAG}_mb_nodense.log | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['dtype'],r['dim'],r['L'],'fwd',round(r['scan_fwd_ms'],4),'bwd',round(r['scan_bwd_ms'],4))"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-300 $O/${TAG}_bench.json
