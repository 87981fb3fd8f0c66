"""times one GEMM shape; run under E3B_GEMM_DEBUG=<mask> to bisect the pipeline (1 skip A load+convert, 2 skip MMA,
4 skip B loads, 8 skip stores)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "equivariant-nn-zoo_b200"))
import torch
from e3b200 import ops
E = 149452
SH = {"s1": (E, 1920, 64), "s2": (E, 64, 1920), "hid": (E, 64, 64), "sc": (150066, 1024, 64), "hid5": (E, 64, 64)}
w = sys.argv[1]
M, N, K = SH[w]
n_prob = 5 if w == "hid5" else 1
probs = []
for _ in range(n_prob):
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
    (Bp,) = ops.gemm_pack([(B, 0, K, 0, 1, 1, 0, N, K)])
    probs.append(ops.gemm_problem(A, Bp, C, M))
for _ in range(3):
    ops.gemm_run(probs)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    ops.gemm_run(probs)
e.record(); torch.cuda.synchronize()
print(w, "dbg", os.environ.get("E3B_GEMM_DEBUG", "0"), "us per launch: %.1f" % (s.elapsed_time(e) / 20 * 1000))
