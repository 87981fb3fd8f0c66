"""cProfile of the host side of one energy+force step (development aid)."""
import cProfile, pstats, os, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "equivariant-nn-zoo_b200")):
    sys.path.insert(0, p)
import torch
import product_harness
from e3_layers.data import Batch, computeEdgeIndex
from e3b200 import synthetic
dev = torch.device("cuda")
model = product_harness.build_product({"config": "config_energy_force", "seed": 0}, torch.float32, dev)
host = synthetic.qm9_like(512, seed=0)
attrs = {"pos": ("node", "1x1o"), "species": ("node", "1x0e"), "_n_nodes": ("graph", "1x0e")}
res = {k: v.to(dev) for k, v in host.items()}
def step():
    batch = Batch(dict(attrs), **{k: v.clone() for k, v in res.items()})
    d, a = computeEdgeIndex(batch.data, batch.attrs, r_max=5.0)
    batch.update(d); batch.attrs.update(a)
    return model(Batch(batch.attrs, **batch.data))
for _ in range(3): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35); print(s.getvalue()[:6000])
