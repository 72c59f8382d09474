#!/bin/bash
# Round-2 first GPU slot: the whole GPU suite (no -x: every family runs), smoke, op-level A/B of every switchable kernel
# against the default and against the reference CUDA kernels (all four stages), bench default (+vs_ref_cuda, cpu arm) and
# the whole-step A/B arms.  Everything lands under gpurun_out/<tag>_*.
TAG=${1:-r2a}
O=gpurun_out
mkdir -p $O
{ nvidia-smi; nproc; free -g; } > $O/${TAG}_env.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python tools/microbench.py --dtypes bf16,f32 --batches 2 --out $O/${TAG}_mb.json > $O/${TAG}_mb.log 2>&1
mb() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python tools/microbench.py --dtypes bf16 --batches 2 --no-ref --out $O/${TAG}_mb_$name.json > $O/${TAG}_mb_$name.log 2>&1
}
mb fwdv2 SMB_FWD_V2=1
mb fwdtma SMB_FWD_V2=2
mb raggv2 SMB_RAGG_V2=1
mb r3v2 SMB_R3_V2=1
mb convv2 SMB_CONV_V2=1 SMB_PERMUTE_V2=1
mb segmin64 SMB_SEG_MIN=64
mb segmin32 SMB_SEG_MIN=32
timeout 900 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
b() { # name, args..., env via B_ENV
  local name=$1; shift
  env $B_ENV timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda "$@" > $O/${TAG}_bench_$name.json 2> $O/${TAG}_bench_$name.err
}
B_ENV="SMB_FUSED_LAYERNORM=1" b ln
B_ENV="SMB_FWD_V2=2 SMB_RAGG_V2=1 SMB_R3_V2=1 SMB_CONV_V2=1 SMB_PERMUTE_V2=1 SMB_FUSED_LAYERNORM=1" b alloptin
B_ENV="SMB_FUSED_LAYERNORM=1" b ln_bf16params --bf16-params
B_ENV="SMB_PAD_CIN=1" b padcin
B_ENV="SMB_DIR_STREAMS=1" b dirstreams
B_ENV="" b graph --cuda-graph
timeout 300 python tools/op_breakdown.py > $O/${TAG}_breakdown.log 2>&1
for f in $O/${TAG}_bench*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print({k:d.get(k) for k in ("value","ms_per_step","native_ms_per_step","host_enqueue_ms_per_step","vs_ref_cuda")}, (d.get("e2e") or {}).get("value"), d.get("cpu_baseline"))
except Exception as e: print("ERR",e)
PY
done
ls -la $O | tail -50
