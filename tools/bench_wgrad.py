#!/usr/bin/env python3
"""Device time of the split-K tcgen05 weight-gradient kernel on the shapes of a W2 training step, next to the
algorithmic HBM bytes (both operands read once) and to torch's matmul on the same operands."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "equivariant-nn-zoo_b200"))
import torch
from e3b200 import ops
dev = torch.device("cuda")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

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

E, N = 149452, 8337
rows = []
for name, R, K1, K2 in [("mlp_last dW  [E,64]^T[E,1920]", E, 64, 1920), ("mlp_hidden dW [E,64]^T[E,64]", E, 64, 64),
                        ("mlp_first dW [E,8]^T[E,64]", E, 8, 64), ("node linear dW [5N,64]^T[5N,64]", 5 * N, 64, 64),
                        ("node linear dW [N,64]^T[N,64]", N, 64, 64)]:
    x = torch.randn(R, K1, device=dev)
    g = torch.randn(R, K2, device=dev)
    t = timeit(lambda: ops.k_wgrad(x, g))
    t_torch = timeit(lambda: x.t() @ g)
    b = R * (K1 + K2) * 4
    rows.append({"shape": name, "ms": t, "torch_mm_ms": t_torch, "alg_GBps": b / t / 1e6, "frac_hbm": b / t / 1e6 / PEAK})
# grouped: the 6 irreps blocks of a node linear map (dims 1,1,3,3,5,5), one launch
mul, dims = 64, [1, 1, 3, 3, 5, 5]
D = sum(d * mul for d in dims)
x = torch.randn(N, D, device=dev); g = torch.randn(N, D, device=dev); W = torch.empty(6 * mul * mul, device=dev)
def grouped():
    probs, off, wo = [], 0, 0
    for d in dims:
        probs.append(ops.wgrad_problem(x, g, W, N * d, mul, mul, a_off=off, a_rows=(D, mul, d), b_off=off, b_rows=(D, mul, d), c_off=wo))
        off += d * mul; wo += mul * mul
    ops.wgrad_run(probs, dev)
t = timeit(grouped)
rows.append({"shape": "grouped 6 irreps blocks of a node linear (one launch)", "ms": t, "alg_GBps": 2 * N * D * 4 / t / 1e6})
# self-connection weights: 6 paths x [N d, 64] x a[N,16] -> W[64,16,64]
a = torch.randn(N, 16, device=dev); Wsc = torch.empty(6 * 64 * 16 * 64, device=dev)
def sc():
    probs, off, wo = [], 0, 0
    for d in dims:
        probs.append(ops.wgrad_problem(x, g, Wsc, N * d, mul, mul, a_off=off, a_rows=(D, mul, d), b_off=off, b_rows=(D, mul, d), aux=a, aux_d=d,
                                       c_off=wo, c_rows=(mul, 16 * mul, mul)))
        off += d * mul; wo += mul * 16 * mul
    ops.wgrad_run(probs, dev)
t = timeit(sc)
rows.append({"shape": "self-connection dW[u,v,w], 6 paths, V=16 (one launch)", "ms": t})
for r in rows:
    print(json.dumps(r))
