#!/usr/bin/env python3
"""Experiment: does running the radial-MLP last layer (writes w [E, 1920]) and the fused TP convolution (reads
w) chunk by chunk -- chunks of whole molecules small enough for w to stay in the 126 MB L2 -- beat the two
whole-batch launches?  (Producer/consumer through L2 instead of HBM; no new kernels.)  Same for the backward
pair: TP backward (writes dw) -> K-long GEMM (reads dw).  Prints one JSON line."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "equivariant-nn-zoo_b200")):
    sys.path.insert(0, p)
import torch

from e3b200 import _lib, ops, plan, synthetic
from e3b200._lib import check, stream

dev = torch.device("cuda")
lib = _lib.load()
host = synthetic.qm9_like(512, seed=0)
pos, n_nodes = host["pos"].to(dev), host["_n_nodes"].reshape(-1).to(dev)
edge_index, n_edges, csr = ops.radius_graph(pos, n_nodes, 5.0)
N, E = pos.shape[0], edge_index.shape[1]
st = plan.with_mul(plan.generated_structures()[3], 64)
P = ops.TPPlan(st)
g = torch.Generator().manual_seed(0)
x = torch.randn(N, P.x_dim, generator=g).to(dev)
Y = torch.randn(E, P.sh_dim, generator=g).to(dev)
h3 = torch.randn(E, 64, generator=g).to(dev)
W3 = torch.randn(64, P.w_dim, generator=g).to(dev)
gy = torch.randn(N, P.y_dim, generator=g).to(dev)
(Bf,) = ops.gemm_pack([(W3, 0, 1, 0, P.w_dim, 1, 0, P.w_dim, 64)])       # fwd: B[n, k] = W3[k, n]
(Bb,) = ops.gemm_pack([(W3, 0, P.w_dim, 0, 1, 1, 0, 64, P.w_dim)])       # bwd: B[n, k] = W3[n, k]
w = torch.empty(E, P.w_dim, device=dev)
gw = torch.empty(E, P.w_dim, device=dev)
gh3 = torch.empty(E, 64, device=dev)
y = torch.empty(N, P.y_dim, device=dev)
gx_edge = torch.empty(E, P.x_dim, device=dev)
gsh = torch.empty(E, P.n_part_f32, P.sh_dim, device=dev)

# chunks of whole molecules: node range [n0, n1) <-> edge-id range [e0, e1) (edges sorted by source)
node_ptr = torch.zeros(n_nodes.numel() + 1, dtype=torch.long)
node_ptr[1:] = n_nodes.cpu().cumsum(0)
row_ptr = csr.in_ptr.cpu()


def chunks(target_edges):
    out, g0 = [], 0
    G = n_nodes.numel()
    while g0 < G:
        g1 = g0 + 1
        while g1 < G and int(row_ptr[node_ptr[g1 + 1]] - row_ptr[node_ptr[g0]]) <= target_edges:
            g1 += 1
        n0, n1 = int(node_ptr[g0]), int(node_ptr[g1])
        out.append((n0, n1, int(row_ptr[n0]), int(row_ptr[n1])))
        g0 = g1
    return out


def fwd_pair(n0, n1, e0, e1):
    ops.gemm_run([ops.gemm_problem(h3, Bf, w, e1 - e0, a_off=e0 * 64, c_off=e0 * P.w_dim, alpha=0.125)])
    check(lib.e3b_tpconv_fwd(P.handle, 0, n1 - n0, E, x.data_ptr(), Y.data_ptr(), w.data_ptr(), csr.in_ptr.data_ptr() + 8 * n0,
                             csr.in_nbr.data_ptr(), csr.in_eid.data_ptr(), y.data_ptr() + 4 * n0 * P.y_dim, stream()))


def bwd_pair(n0, n1, e0, e1):
    check(lib.e3b_tpconv_bwd(P.handle, 0, n1 - n0, E, x.data_ptr(), Y.data_ptr(), w.data_ptr(), gy.data_ptr() + 4 * n0 * P.y_dim,
                             csr.in_ptr.data_ptr() + 8 * n0, csr.in_nbr.data_ptr(), csr.in_eid.data_ptr(),
                             gx_edge.data_ptr(), gsh.data_ptr(), gw.data_ptr(), stream()))
    ops.gemm_run([ops.gemm_problem(gw, Bb, gh3, e1 - e0, a_off=e0 * P.w_dim, c_off=e0 * 64, alpha=0.125)])


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


res = {"N": N, "E": E, "w_GB": E * P.w_dim * 4 / 1e9}
fwd_pair(0, N, 0, E)
y_ref = y.clone()
bwd_pair(0, N, 0, E)
gh_ref, gw_ref = gh3.clone(), gw.clone()
for target in (None, 64000, 32000, 16000, 8000, 4000):
    cs = [(0, N, 0, E)] if target is None else chunks(target)
    y.zero_(), gh3.zero_()

    def run_f():
        for c in cs:
            fwd_pair(*c)

    def run_b():
        for c in cs:
            bwd_pair(*c)

    # replay as CUDA graphs so that launch overhead of the many small launches does not decide the result
    def graphed(fn):
        fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        return gr.replay

    tf, tb = timeit(graphed(run_f)), timeit(graphed(run_b))
    ok = bool(torch.equal(y, y_ref) and torch.equal(gh3, gh_ref) and torch.equal(gw, gw_ref))
    res[str(target)] = {"chunks": len(cs), "fwd_pair_ms": tf, "bwd_pair_ms": tb, "identical": ok}
print(json.dumps(res))
