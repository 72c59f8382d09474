"""Timing of the fused instance-norm forward / backward at the model's large shapes (CUDA events, L2 flushed).  Run once per
library variant: SMB_LIB=segmamba_b200/variants/lib_X.so python tools/norm_bench.py"""
import json, os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmamba_b200 import _lib
from segmamba_b200.instance_norm import fused_instance_norm

flush = torch.empty(128 * 1024 * 1024, device="cuda")
res = {}
for shape in ((2, 48, 128, 128, 128), (2, 96, 64, 64, 64), (2, 48, 64, 64, 64)):
    x = torch.randn(shape, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last_3d).requires_grad_()
    r = torch.randn_like(x).requires_grad_()
    dy = torch.randn_like(x)
    for mode, kw in (("plain", {}), ("residual", dict(add=r)), ("two_norms", dict(add=r, add_norm=True))):
        def run():
            y = fused_instance_norm(x, "leaky_relu", 0.01, **kw)
            y.backward(dy)
            x.grad = None; r.grad = None
        for _ in range(2):
            run()
        tf, tb = [], []
        for _ in range(6):
            flush.fill_(1.0)
            with _lib.profile() as prof:
                run()
                d = prof.durations()
            tf.append(sum(v[0] for (op, m), v in d.items() if op == "instnorm_fwd"))
            tb.append(sum(v[0] for (op, m), v in d.items() if op == "instnorm_bwd"))
        res[f"{shape[1]}x{shape[2]}^3 {mode}"] = (round(statistics.median(tf), 4), round(statistics.median(tb), 4))
print(os.environ.get("SMB_LIB", "default"), json.dumps(res))
