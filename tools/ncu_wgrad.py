"""one weight-gradient shape, a few launches (for ncu --set full)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "equivariant-nn-zoo_b200"))
import torch
from e3b200 import ops
R, K1, K2 = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (149452, 64, 1920)
x = torch.randn(R, K1, device="cuda"); g = torch.randn(R, K2, device="cuda")
for _ in range(4):
    ops.k_wgrad(x, g)
torch.cuda.synchronize()
