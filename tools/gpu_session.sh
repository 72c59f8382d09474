#!/bin/bash
# One gpurun call = validation + A/B measurements + profiles, everything under gpurun_out/<tag>_*.
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_session.sh r2a'
#
# Stages can be switched off with SKIP="tests ncu ..." (space separated: tests smoke mb bench ab sanitizer variants ncu ncufull).
# Numbers printed under ncu are never bench values; the bench lines come from the plain runs.
TAG=${1:-s}
O=gpurun_out
mkdir -p $O
skip() { [[ " $SKIP " == *" $1 "* ]]; }
{ nvidia-smi; nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv; nproc; } > $O/${TAG}_env.txt 2>&1

if ! skip tests; then
  timeout 900 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
fi
if ! skip smoke; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1
fi
if ! skip mb; then
  timeout 600 python tools/microbench.py --dtypes bf16,f32 --batches 2 --out $O/${TAG}_mb.json > $O/${TAG}_mb.log 2>&1
fi
if ! skip bench; then
  timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
  timeout 600 python bench.py --steps 10 --warmup 3 --cuda-graph --no-cpu-baseline > $O/${TAG}_bench_graph.json 2> $O/${TAG}_bench_graph.err
fi
if ! skip ab; then   # A/B of the opt-in paths against the default, same box, back to back
  SMB_FUSED_LAYERNORM=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_ln.json 2> $O/${TAG}_bench_ln.err
  SMB_DIR_STREAMS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_dirstreams.json 2> $O/${TAG}_bench_dirstreams.err
  SMB_DIR_STREAMS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --cuda-graph > $O/${TAG}_bench_dirstreams_graph.json 2> $O/${TAG}_bench_dirstreams_graph.err
  SMB_FWD_V2=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_fwdv2.json 2> $O/${TAG}_bench_fwdv2.err
  SMB_CONV_V2=1 timeout 300 python tools/microbench.py --dtypes bf16 --batches 2 --stages 0,1 --no-ref --out $O/${TAG}_mb_convv2.json > $O/${TAG}_mb_convv2.log 2>&1
  SMB_FWD_V2=1 SMB_RAGG_V2=1 SMB_CONV_V2=1 SMB_FUSED_LAYERNORM=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_all_optin.json 2> $O/${TAG}_bench_all_optin.err
  SMB_FWD_V2=1 SMB_RAGG_V2=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_fwdv2_raggv2.json 2> $O/${TAG}_bench_fwdv2_raggv2.err
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --bf16-params > $O/${TAG}_bench_bf16params.json 2> $O/${TAG}_bench_bf16params.err
  SMB_PAD_CIN=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_padcin.json 2> $O/${TAG}_bench_padcin.err
  SMB_CUDNN_BENCH_LIMIT=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_cudnnall.json 2> $O/${TAG}_bench_cudnnall.err
  SMB_PERMUTE_V2=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_permutev2.json 2> $O/${TAG}_bench_permutev2.err
  SMB_R3_V2=1 timeout 300 python tools/microbench.py --dtypes bf16,f32 --batches 2 --stages 0,1 --no-ref --out $O/${TAG}_mb_r3v2.json > $O/${TAG}_mb_r3v2.log 2>&1
  SMB_R3_V2=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_r3v2.json 2> $O/${TAG}_bench_r3v2.err
  SMB_RAGG_V2=1 timeout 300 python tools/microbench.py --dtypes bf16 --batches 2 --stages 0,1 --no-ref --out $O/${TAG}_mb_raggv2.json > $O/${TAG}_mb_raggv2.log 2>&1
  SMB_FWD_V2=2 timeout 300 python tools/microbench.py --dtypes bf16 --batches 2 --stages 0,1 --no-ref --out $O/${TAG}_mb_fwdtma.json > $O/${TAG}_mb_fwdtma.log 2>&1
  SMB_FWD_V2=1 timeout 300 python tools/microbench.py --dtypes bf16 --batches 2 --stages 0,1 --no-ref --out $O/${TAG}_mb_fwdv2.json > $O/${TAG}_mb_fwdv2.log 2>&1
  for s in 32 64 128; do
    SMB_SEG_MIN=$s timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_segmin$s.json 2> $O/${TAG}_bench_segmin$s.err
  done
  SMB_FUSED_LAYERNORM=1 timeout 300 python -m pytest tests/test_gpu_zz_layernorm.py -q > $O/${TAG}_pytest_ln.log 2>&1
  timeout 300 python tools/op_breakdown.py > $O/${TAG}_breakdown.log 2>&1
  timeout 900 python tools/ref_equivalent_step.py --native --steps 5 --out $O/${TAG}_ref_equivalent_step.json > $O/${TAG}_ref_equivalent_step.log 2>&1
fi
if ! skip sanitizer; then   # memcheck + racecheck of the opt-in kernels at small sizes (their first hardware runs)
  cat > $O/${TAG}_san.py <<'PY'
import os, sys, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import rand_scan_inputs
from segmamba_b200 import selective_scan_cuda as ssc, causal_conv1d_cuda as cc
from segmamba_b200.layer_norm import fused_layer_norm
for mode in ("1", "2"):
    os.environ["SMB_FWD_V2"] = mode; os.environ["SMB_RAGG_V2"] = "1"; os.environ["SMB_CONV_V2"] = "1"; os.environ["SMB_R3_V2"] = "1"; os.environ["SMB_PERMUTE_V2"] = "1"
    for direction in (0, 1):
        d = rand_scan_inputs(3, 2, 40, 1000, 16, 1, torch.bfloat16)
        B, C = d["B"].unsqueeze(1), d["C"].unsqueeze(1)
        o = ssc.fwd_ex(d["u"], d["delta"], d["A"], B, C, d["D"], d["z"], d["delta_bias"], True, direction=direction, want_hstates=True)
        ssc.bwd_ex(d["u"], d["delta"], d["A"], B, C, d["D"], d["z"], d["delta_bias"], d["dout"], None, True, False, direction=direction, hstates=o[3])
        x = torch.randn(2, 40, 1000, device="cuda").bfloat16(); w = torch.randn(40, 4, device="cuda"); b = torch.randn(40, device="cuda")
        y = cc.causal_conv1d_fwd_ex(x, w, b, True, direction=direction)
        cc.causal_conv1d_bwd_ex(x, w, b, y, None, True, direction=direction)
        cc.seq_permute(cc.seq_permute(x, 10), 10, inverse=True)
t = torch.randn(700, 96, device="cuda").bfloat16().requires_grad_(); g = torch.rand(96, device="cuda").requires_grad_(); bb = torch.zeros(96, device="cuda").requires_grad_()
fused_layer_norm(t, g, bb).float().sum().backward()
torch.cuda.synchronize(); print("sanitizer workload done")
PY
  timeout 900 compute-sanitizer --tool memcheck python $O/${TAG}_san.py > $O/${TAG}_san_memcheck.log 2>&1
  timeout 900 compute-sanitizer --tool racecheck python $O/${TAG}_san.py > $O/${TAG}_san_racecheck.log 2>&1
fi
if ! skip variants; then   # build-time tuning variants, if tools/build_variants.py was run before the call
  for lib in segmamba_b200/variants/lib_*.so; do
    [ -e "$lib" ] || continue
    v=$(basename $lib .so)
    SMB_LIB=$PWD/$lib timeout 300 python tools/microbench.py --dtypes bf16 --batches 2 --stages 0,1 --no-ref --out $O/${TAG}_mb_$v.json > $O/${TAG}_mb_$v.log 2>&1
    SMB_LIB=$PWD/$lib timeout 300 python tools/op_breakdown.py > $O/${TAG}_breakdown_$v.log 2>&1
    case $v in lib_poly*) SMB_LIB=$PWD/$lib SMB_FWD_V2=1 SMB_RAGG_V2=1 timeout 300 python tools/microbench.py --dtypes bf16 --batches 2 --stages 0,1 --no-ref --out $O/${TAG}_mb_${v}_fwdv2.json > $O/${TAG}_mb_${v}_fwdv2.log 2>&1;; esac
  done
fi
if ! skip ncu; then  # launch list of one training step (kernel shares)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${TAG}_launches_step.csv \
      python tools/profile_step.py --what step > $O/${TAG}_ncu_step.log 2>&1
  python tools/launch_summary.py $O/${TAG}_launches_step.csv > $O/${TAG}_launches_step_summary.txt 2>&1
fi
if ! skip ncufull; then  # one full capture of the scan kernels (stage 0, training batch)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_ -c 8 -o $O/${TAG}_scan_full -f \
      python tools/profile_step.py --what scan > $O/${TAG}_ncu_scan.log 2>&1
  ncu -i $O/${TAG}_scan_full.ncu-rep --page raw --csv > $O/${TAG}_scan_raw.csv 2>/dev/null
  python tools/ncu_raw_summary.py $O/${TAG}_scan_raw.csv > $O/${TAG}_scan_summary.txt 2>&1
fi
python tools/session_report.py $TAG $O > $O/${TAG}_report.txt 2>&1
cat $O/${TAG}_report.txt | cut -c1-250
ls -la $O | tail -40
