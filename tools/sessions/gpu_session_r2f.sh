#!/bin/bash
TAG=${1:-r2f}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -p no:cacheprovider > $O/${TAG}_pytest_gemm.log 2>&1; echo "gemm pytest rc=$?"; grep -n "smb_gemm:\|rel err" $O/${TAG}_pytest_gemm.log | head -8 | cut -c1-300; tail -4 $O/${TAG}_pytest_gemm.log | cut -c1-200
timeout 300 python tools/gemm_bench.py --out $O/${TAG}_gemm_bench.json 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); print(r['stage'], r['op'].ljust(38), r['M'],r['N'],r['K'], 'native %.4f lib %.4f  x%.2f' % (r['native_ms'], r['library_ms'], r['ratio_lib_over_native']))"
timeout 600 python -m pytest tests/test_gpu_scan.py tests/test_gpu_scan_variants.py -q -x -p no:cacheprovider > $O/${TAG}_pytest_scan.log 2>&1; echo "scan pytest rc=$?"; tail -4 $O/${TAG}_pytest_scan.log | cut -c1-200
for opc in 0 1; do
echo "== SMB_R3_OPC=$opc (0 = auto)"
SMB_R3_OPC=$opc timeout 200 python tools/microbench.py --dtypes bf16 --batches 2 --no-ref --out $O/${TAG}_mb_opc$opc.json 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); print(r['dim'],r['L'],'fwd',round(r['scan_fwd_ms'],4),'bwd',round(r['scan_bwd_ms'],4))"
done
