"""Timing of the other BASELINE workloads on one GPU (W4 small-molecule score evaluation, W5 protein-sized
C-alpha graph); bench.py measures the headline W2.  Prints one JSON line per workload."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "equivariant-nn-zoo_b200")):
    sys.path.insert(0, p)
import torch

import product_harness
from e3b200 import synthetic

dev = torch.device("cuda")


def timeit(fn, reps):
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


def run(name, meta, inputs, reps, **kw):
    model = product_harness.build_product(meta, torch.float32, dev)
    dev_in = {k: v.to(dev) for k, v in inputs.items()}
    ei = dev_in.pop("edge_index", None)
    with torch.no_grad():
        ms = timeit(lambda: product_harness.run_product(model, dev_in, torch.float32, dev, edge_index=ei, **kw), reps)
    n = inputs[next(k for k in ("pos", "CA") if k in inputs)].shape[0]
    E = ei.shape[1] if ei is not None else None
    print(json.dumps({"workload": name, "atoms": n, "edges": E, "ms_per_evaluation": ms, "atoms_per_s": n / ms * 1e3,
                      "edges_per_s": (E / ms * 1e3) if E else None, "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}),
          flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["W4", "W5"]
    if "W4" in which:
        run("W4 config_diffusion score evaluation (no-grad), 128 molecules, complete graphs",
            {"config": "config_diffusion", "seed": 0}, synthetic.diffusion_like(128, seed=0), 20)
    if "W5" in which:
        for n_res in (2000, 4000):
            torch.cuda.reset_peak_memory_stats()
            run(f"W5 config_diffusion_CA score evaluation (no-grad), one graph of {n_res} residues",
                {"config": "config_diffusion_CA", "seed": 0}, synthetic.protein_like(n_res, seed=0), 5)
