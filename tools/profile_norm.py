"""Run the fused instance-norm fwd+bwd at the largest shape inside a profiler range (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmamba_b200.instance_norm import fused_instance_norm

x = torch.randn(2, 48, 128, 128, 128, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last_3d).requires_grad_()
r = torch.randn_like(x).requires_grad_()
dy = torch.randn_like(x)


def run():
    y = fused_instance_norm(x, "leaky_relu", 0.01)
    y.backward(dy)
    y2 = fused_instance_norm(x, "leaky_relu", 0.01, add=r, add_norm=True)
    y2.backward(dy)


for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
