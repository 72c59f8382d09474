#!/bin/bash
# Last slot: the driver's commands on the final tree + one R3 octet-walk sweep
TAG=${1:-r2y}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${TAG}_smoke.log | cut -c1-250
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-260 $O/${TAG}_bench.json
for opc in 2 4 6; do
echo "== SMB_R3_OPC=$opc"
SMB_R3_OPC=$opc timeout 120 python tools/microbench.py --dtypes bf16 --batches 2 --stages 0,1,2 --no-ref --iters 6 --out $O/${TAG}_mb_opc$opc.json 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); print(r['dim'],r['L'],'bwd',round(r['scan_bwd_ms'],4))"
done
