"""torch.profiler kernel table of one energy+force step (W2) -- development aid for finding where
the step time goes; the judged evidence is the ncu launch list / full capture under profiles/."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "equivariant-nn-zoo_b200")):
    sys.path.insert(0, p)
import torch
from torch.profiler import ProfilerActivity, profile

import product_harness
from e3_layers.data import Batch, computeEdgeIndex
from e3b200 import synthetic

G = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda")
model = product_harness.build_product({"config": "config_energy_force", "seed": 0}, torch.float32, dev)
host = synthetic.qm9_like(G, seed=0)
attrs = {"pos": ("node", "1x1o"), "species": ("node", "1x0e"), "_n_nodes": ("graph", "1x0e")}
res = {k: v.to(dev) for k, v in host.items()}


def step():
    batch = Batch(dict(attrs), **{k: v.clone() for k, v in res.items()})
    d, a = computeEdgeIndex(batch.data, batch.attrs, r_max=5.0)
    batch.update(d)
    batch.attrs.update(a)
    return model(Batch(batch.attrs, **batch.data))


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
