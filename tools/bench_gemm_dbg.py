"""GEMM bisection with E3B_GEMM_DEBUG (development aid): rotates 4 buffer sets so operands come from HBM."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "equivariant-nn-zoo_b200")):
    sys.path.insert(0, p)
import torch
from e3b200 import ops
dev = torch.device("cuda"); E = 149452
SH = {"s1": (E, 1920, 64, 0), "s2": (E, 64, 1920, 0), "hid": (E, 64, 64, 2), "hidb": (E, 64, 64, 3), "first": (E, 64, 8, 2)}
for name in (sys.argv[1:] or ["s1", "s2"]):
    M, N, K, epi = SH[name]
    sets = []
    for _ in range(4 if M * max(N, K) < 3e7 else 1):
        A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); C = torch.empty(M, N, device=dev)
        H = torch.rand(M, N, device=dev) if epi == 3 else None
        (Bp,) = ops.gemm_pack([(B, 0, K, 0, 1, 1, 0, N, K)])
        sets.append([ops.gemm_problem(A, Bp, C, M, epilogue=epi, H=H, act_cst=1.8782, alpha=0.125)])
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    for i in range(3): ops.gemm_run(sets[i % len(sets)])
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(12):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.gemm_run(sets[i % len(sets)]); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    print(os.environ.get("E3B_GEMM_DEBUG", "0"), name, f"{tot/12*1e3:.1f} us", flush=True)
