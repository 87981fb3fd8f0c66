import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "equivariant-nn-zoo_b200")):
    sys.path.insert(0, p)
import torch
from e3b200 import ops
dev = torch.device("cuda"); E = 149452
for name, M, N, K in [("s1", E, 1920, 64), ("s2", E, 64, 1920)]:
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); C = torch.empty(M, N, device=dev)
    (Bp,) = ops.gemm_pack([(B, 0, K, 0, 1, 1, 0, N, K)])
    prob = [ops.gemm_problem(A, Bp, C, M)]
    for _ in range(3): ops.gemm_run(prob)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): ops.gemm_run(prob)
    e.record(); torch.cuda.synchronize()
    print(os.environ.get("E3B_GEMM_DEBUG", "0"), name, f"{s.elapsed_time(e)/10*1e3:.1f} us", flush=True)
