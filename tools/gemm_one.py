import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segmamba_b200 import gemm as G
bf = torch.bfloat16
T = 524288
a, b = torch.randn(192, 48, device="cuda").to(bf), torch.randn(T, 48, device="cuda").to(bf)
for _ in range(3):
    G.gemm(a, b)
torch.cuda.synchronize()
