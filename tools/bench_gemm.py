"""Micro-benchmark of the tcgen05 3xTF32 GEMM on the shapes of the interaction block (development
aid; CUDA events, L2 flushed between launches by the operands' own size where they exceed it)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "equivariant-nn-zoo_b200")):
    sys.path.insert(0, p)
import torch

from e3b200 import ops

dev = torch.device("cuda")
E, Nn = 149452, 8337
SHAPES = [
    ("radial last fwd  [E,64]x[64,1920]", E, 1920, 64),
    ("radial last bwd  [E,1920]x[1920,64]", E, 64, 1920),
    ("radial hidden    [E,64]x[64,64]", E, 64, 64),
    ("radial first     [E,8]x[8,64]", E, 64, 8),
    ("linear block l=2 [5N,64]x[64,64]", 5 * Nn, 64, 64),
    ("post-TP l=1      [3N,384]x[384,64]", 3 * Nn, 64, 384),
    ("sc fwd l=2       [5N,64]x[64,1024]", 5 * Nn, 1024, 64),
]


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


for name, M, N, K in SHAPES:
    A = torch.randn(M, K, device=dev)
    B = torch.randn(N, K, device=dev)
    C = torch.empty(M, N, device=dev)
    (Bp,) = ops.gemm_pack([(B, 0, K, 0, 1, 1, 0, N, K)])
    prob = [ops.gemm_problem(A, Bp, C, M)]
    t = timeit(lambda: ops.gemm_run(prob))
    torch.backends.cuda.matmul.allow_tf32 = False
    t_ref = timeit(lambda: torch.matmul(A, B.T, out=C))
    gb = (M * K + N * K + M * N) * 4 / 1e9
    print(f"{name:40s} ours {t*1e3:8.1f} us  {2*M*N*K/t/1e9:8.1f} TFLOP/s(fp32-equiv) {gb/t*1e3:7.1f} GB/s | "
          f"torch fp32 {t_ref*1e3:8.1f} us", flush=True)
