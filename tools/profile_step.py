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

TRAIN = "--train" in sys.argv          # force-matching training step (second-order mode) instead of the evaluation
args = [a for a in sys.argv[1:] if not a.startswith("--")]
G = int(args[0]) if args else 512
dev = torch.device("cuda")
model = product_harness.build_product({"config": "config_energy_force", "seed": 0}, torch.float32, dev)
host = synthetic.qm9_like(G, seed=0)
attrs = {"pos": ("node", "1x1o"), "species": ("node", "1x0e"), "_n_nodes": ("graph", "1x0e")}
res = {k: v.to(dev) for k, v in host.items()}
if TRAIN:
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)


def step():
    batch = Batch(dict(attrs), **{k: v.clone() for k, v in res.items()})
    d, a = computeEdgeIndex(batch.data, batch.attrs, r_max=5.0)
    batch.update(d)
    batch.attrs.update(a)
    out = model(Batch(batch.attrs, **batch.data))
    if TRAIN:
        loss = 1e3 * (out["energy"] ** 2).mean() + 3e4 * ((out["forces"] - 0.1) ** 2).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    return out


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
if "--cpu" in sys.argv:
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=40, max_name_column_width=70))
