"""ncu target: a few launches of the tcgen05 GEMM on the interaction-block shapes (development aid)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "equivariant-nn-zoo_b200")):
    sys.path.insert(0, p)
import torch

from e3b200 import ops

dev = torch.device("cuda")
E = 149452
which = sys.argv[1:] or ["s1", "s2", "hid"]
SH = {"s1": (E, 1920, 64), "s2": (E, 64, 1920), "hid": (E, 64, 64), "sc": (41685, 1024, 64)}
for w in which:
    M, N, K = SH[w]
    A = torch.randn(M, K, device=dev)
    B = torch.randn(N, K, device=dev)
    C = torch.empty(M, N, device=dev)
    (Bp,) = ops.gemm_pack([(B, 0, K, 0, 1, 1, 0, N, K)])
    prob = [ops.gemm_problem(A, Bp, C, M)]
    for _ in range(3):
        ops.gemm_run(prob)
    torch.cuda.synchronize()
