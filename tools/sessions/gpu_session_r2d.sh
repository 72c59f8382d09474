#!/bin/bash
TAG=${1:-r2d}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -p no:cacheprovider > $O/${TAG}_pytest_gemm.log 2>&1; echo "gemm pytest rc=$?"; grep -n "smb_gemm:" $O/${TAG}_pytest_gemm.log | head -5 | cut -c1-300; tail -4 $O/${TAG}_pytest_gemm.log | cut -c1-200
for dbg in 0 1 3 7; do
  echo "== SMB_R3_DBG=$dbg"
  SMB_R3_DBG=$dbg timeout 200 python tools/microbench.py --dtypes bf16 --batches 2 --stages 0,1 --no-ref --out $O/${TAG}_mb_dbg$dbg.json 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); print(r['dim'],r['L'],'fwd',round(r['scan_fwd_ms'],4),'bwd',round(r['scan_bwd_ms'],4))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_bwd -c 6 -o $O/${TAG}_scan_bwd_full -f python tools/profile_step.py --what scan > $O/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i $O/${TAG}_scan_bwd_full.ncu-rep --page raw --csv > $O/${TAG}_scan_bwd_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py $O/${TAG}_scan_bwd_raw.csv > $O/${TAG}_scan_bwd_summary.txt 2>&1; cat $O/${TAG}_scan_bwd_summary.txt | cut -c1-250
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_sliding_window.py -q -p no:cacheprovider > $O/${TAG}_pytest_model.log 2>&1; echo "model pytest rc=$?"; tail -5 $O/${TAG}_pytest_model.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-260 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
SMB_CUDNN_BENCH_LIMIT=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda > $O/${TAG}_bench_cudnnall.json 2> $O/${TAG}_bench_cudnnall.err; cut -c1-260 $O/${TAG}_bench_cudnnall.json
