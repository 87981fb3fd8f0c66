#!/usr/bin/env python3
"""W4 (BASELINE config #4): batched score evaluation over sampler steps.  Times predictor-corrector
iterations (2 score evaluations + position updates each) of config_diffusion on 128 synthetic molecules with
complete graphs, eager loop vs one captured CUDA graph replayed per iteration.  One JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "equivariant-nn-zoo_b200")):
    sys.path.insert(0, p)
import torch

import harness
import product_harness
from e3_layers.data import Batch
from e3_layers.run import VPSDE, EulerMaruyamaPredictor, LangevinCorrector, get_pc_sampler
from e3b200 import synthetic

dev = torch.device("cuda")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 400
inputs = synthetic.diffusion_like(G, seed=0)
model = product_harness.build_product({"config": "config_diffusion", "seed": 0}, torch.float32, dev)
data = harness.cast_inputs(inputs, torch.float32, dev)
res = {}


def run(graph, n_iter):
    sde = VPSDE({"pos": 3}, N=1000)
    sampler = get_pc_sampler(sde, EulerMaruyamaPredictor, LangevinCorrector, lambda b: b, snr=0.16, max_iterations=n_iter, graph=graph)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, nfe = sampler(model, Batch(harness.attrs_for(data), **data))
    torch.cuda.synchronize()
    assert torch.isfinite(out["pos"]).all()
    return time.perf_counter() - t0


for name, graph in (("eager", False), ("cuda_graph", True)):
    run(graph, 10)                                                        # warm-up (plans, weight packs)
    short, full = run(graph, iters // 4), run(graph, iters)
    per_iter = (full - short) / (iters - iters // 4)                      # steady state: set-up / capture cost cancels
    res[name] = {"ms_per_iteration": per_iter * 1e3, "score_evaluations_per_s": 2 / per_iter,
                 "atoms_per_s": inputs["pos"].shape[0] * 2 / per_iter, "whole_run_s": full,
                 "setup_or_capture_s": full - per_iter * iters}
print(json.dumps({"workload": f"W4 config_diffusion PC sampler, {G} molecules ({inputs['pos'].shape[0]} atoms, "
                              f"{inputs['edge_index'].shape[1]} edges), {iters} iterations (2 score evaluations each)", **res}))
