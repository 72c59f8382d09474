#!/bin/bash
# Final slot of the round: what the driver runs (GPU suite, smoke, bench) + the records DESIGN.md cites
TAG=${1:-r2z}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${TAG}_pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${TAG}_smoke.log | cut -c1-250
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-260 $O/${TAG}_bench.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_reference_arm.err; echo "ref arm rc=$?"; cut -c1-200 $O/${TAG}_bench_reference_arm.json
timeout 600 python tools/microbench.py --dtypes bf16,f32 --batches 2 --out $O/${TAG}_microbench_vs_refcuda.json > $O/${TAG}_mb.log 2>&1; grep -c scan_fwd_ms $O/${TAG}_mb.log
timeout 300 python tools/op_breakdown.py > $O/${TAG}_op_breakdown.txt 2>&1; head -12 $O/${TAG}_op_breakdown.txt | cut -c1-150
