#!/bin/bash
TAG=${1:-r2l}
O=gpurun_out
mkdir -p $O
# launch list of ONE training step (kernel shares), profile range = the 4th step
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_step.csv python tools/profile_step.py --what step > $O/${TAG}_ncu_step.log 2>&1; echo "ncu step rc=$?"
python tools/launch_summary.py $O/${TAG}_launches_step.csv 60 detail > $O/${TAG}_launches_step_summary.txt 2>&1; head -70 $O/${TAG}_launches_step_summary.txt | cut -c1-170
# full capture of the scan kernels (stage 0, training batch): forward agg + main, R1, R3
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:scan_ -o $O/${TAG}_scan_full -f python tools/profile_step.py --what scan > $O/${TAG}_ncu_scan.log 2>&1; echo "ncu scan rc=$?"
ncu -i $O/${TAG}_scan_full.ncu-rep --page raw --csv > $O/${TAG}_scan_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py $O/${TAG}_scan_raw.csv > $O/${TAG}_scan_summary.txt 2>&1; cat $O/${TAG}_scan_summary.txt | cut -c1-220
ncu -i $O/${TAG}_scan_full.ncu-rep --page source --csv --kernel-name regex:main2 > $O/${TAG}_r3_source.csv 2>/dev/null
# side streams A/B with today's kernels, eager and inside the graph
SMB_DIR_STREAMS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda > $O/${TAG}_bench_streams.json 2> $O/${TAG}_bench_streams.err; cut -c1-230 $O/${TAG}_bench_streams.json
SMB_DIR_STREAMS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda --cuda-graph > $O/${TAG}_bench_streams_graph.json 2> $O/${TAG}_bench_streams_graph.err; cut -c1-230 $O/${TAG}_bench_streams_graph.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda --bf16-params > $O/${TAG}_bench_bf16params.json 2> $O/${TAG}_bench_bf16params.err; cut -c1-230 $O/${TAG}_bench_bf16params.json
