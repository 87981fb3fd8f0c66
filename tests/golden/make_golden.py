"""Generate the golden fixtures under tests/golden/ by running the GENUINE reference package
(/root/reference/e3_layers, unmodified) in the build container.

The reference's third-party imports (e3nn 0.4.4, ml_collections, torch_runstats, ase, h5py)
are not installable here; ``oracle/shims.py`` routes them to the restatement in
``oracle/e3nn_ops.py``.  The fixtures therefore pin the reference's OWN code (config builders,
layer order, irreps wiring, key mapping, neighbour list, embeddings, heads, autograd forces)
while the e3nn arithmetic underneath stays "parity unpinned" (see oracle/__init__.py).

Weights are not stored: ``reseed_parameters`` (tests/param_init.py) overwrites every parameter
deterministically from its state_dict NAME, so the test side can rebuild identical weights for
the oracle and for the CUDA product at full reference width.

Run (build container only):  python tests/golden/make_golden.py
Deviation from "unmodified": in fp64 runs ``OneHotEncoding.forward`` is wrapped to cast its
float32 one-hot (reference defect D6, nn/embedding.py:275-277) to float64.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "equivariant-nn-zoo_b200"))

from e3b200 import synthetic  # noqa: E402  (plain torch helpers, no CUDA needed)
from oracle import shims  # noqa: E402
from param_init import reseed_parameters  # noqa: E402

ref = shims.import_reference()
from e3_layers import configs  # noqa: E402
from e3_layers.data import Batch, computeEdgeIndex  # noqa: E402
from e3_layers.utils import build  # noqa: E402
import e3_layers.nn.embedding as ref_embedding  # noqa: E402


def _patch_onehot_fp64():
    orig = ref_embedding.OneHotEncoding.forward

    def fwd(self, data, attrs):
        d, a = orig(self, data, attrs)
        return {k: v.to(torch.get_default_dtype()) for k, v in d.items()}, a

    ref_embedding.OneHotEncoding.forward = fwd
    return orig


def run_model(cfg_fn, spec, inputs, attrs, seed, dtype, out_keys, pre_edge=None, torch_seed=0):
    torch.set_default_dtype(dtype)
    orig = _patch_onehot_fp64() if dtype == torch.float64 else None
    try:
        cfg = cfg_fn(spec) if spec is not None else cfg_fn()
        model = build(cfg.model_config)
        reseed_parameters(model, seed)
        model.eval()
        tensors = {k: (v.to(dtype) if v.is_floating_point() else v.clone()) for k, v in inputs.items()}
        batch = Batch(dict(attrs), **tensors)
        if pre_edge is not None:  # dataset-style preprocess (compute_edge as a function, D1)
            d, a = computeEdgeIndex(batch.data, batch.attrs, **pre_edge)
            batch.update(d)
            batch.attrs.update(a)
            batch = Batch(batch.attrs, **batch.data)
        torch.manual_seed(torch_seed)
        out = model(batch)
        res = {k: out[k].detach().cpu().numpy() for k in out_keys}
        res["edge_index"] = out["edge_index"].cpu().numpy()
        return res
    finally:
        if orig is not None:
            ref_embedding.OneHotEncoding.forward = orig
        torch.set_default_dtype(torch.float32)


def save(name, meta, inputs, out32, out64):
    arrs = {}
    for k, v in inputs.items():
        arrs["in/" + k] = v.numpy()
    for k, v in out32.items():
        if k != "node_features":  # kept in fp64 only (fixture size)
            arrs["out32/" + k] = v
    for k, v in out64.items():
        arrs["out64/" + k] = v
    arrs["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(name, {k: v.shape for k, v in arrs.items() if k != "meta"}, os.path.getsize(path), "bytes")


NODE = lambda ir: ("node", ir)  # noqa: E731


def model_cases():
    attrs_mol = {"pos": NODE("1x1o"), "species": NODE("1x0e"), "_n_nodes": ("graph", "1x0e")}

    # 1. config_energy_force at full reference width (n_dim 64, l_max 2, 5 layers, r_max 5)
    inp = synthetic.qm9_like(3, seed=11, n_min=3, n_max=9)
    kw = dict(inputs=inp, attrs=attrs_mol, seed=1, out_keys=["energy", "forces", "node_features"],
              pre_edge={"r_max": 5.0})
    save("model_energy_force", {"config": "config_energy_force", "seed": 1, "r_max": 5.0},
         inp, run_model(configs.config_energy_force, None, dtype=torch.float32, **kw),
         run_model(configs.config_energy_force, None, dtype=torch.float64, **kw))

    # 2. config_energy (l_max 3 features -> 48-path layers, energy only, r_max 4)
    inp = synthetic.qm9_like(2, seed=12, n_min=4, n_max=8)
    kw = dict(inputs=inp, attrs=attrs_mol, seed=2, out_keys=["total_energy", "node_features"],
              pre_edge={"r_max": 4.0})
    save("model_energy", {"config": "config_energy", "seed": 2, "r_max": 4.0},
         inp, run_model(configs.config_energy, None, dtype=torch.float32, **kw),
         run_model(configs.config_energy, None, dtype=torch.float64, **kw))

    # 3. config_dipole (1x1o head)
    inp = synthetic.qm9_like(3, seed=13, n_min=3, n_max=8, species_choices=tuple(range(1, 18)))
    kw = dict(inputs=inp, attrs=attrs_mol, seed=3, out_keys=["dipole", "node_features"], pre_edge={"r_max": 5.0})
    save("model_dipole", {"config": "config_dipole", "seed": 3, "r_max": 5.0},
         inp, run_model(configs.config_dipole, None, dtype=torch.float32, **kw),
         run_model(configs.config_dipole, None, dtype=torch.float64, **kw))

    # 4. config_diffusion (score head; bond one-hot + time embedding; complete graphs)
    inp = synthetic.diffusion_like(3, seed=14, n_min=3, n_max=7)
    attrs = dict(attrs_mol, t=("graph", "1x0e"), bond_type=("edge", "1x0e"), _n_edges=("graph", "1x0e"))
    kw = dict(inputs=inp, attrs=attrs, seed=4, out_keys=["score", "node_features"])
    save("model_diffusion", {"config": "config_diffusion", "seed": 4, "spec": ""},
         inp, run_model(configs.config_diffusion, "", dtype=torch.float32, **kw),
         run_model(configs.config_diffusion, "", dtype=torch.float64, **kw))

    # 4b. config_diffusion 'nll' spec: score = +d(nll)/d(pos) through GradientOutput
    kw = dict(inputs=inp, attrs=attrs, seed=5, out_keys=["score", "nll"])
    save("model_diffusion_nll", {"config": "config_diffusion", "seed": 5, "spec": "nll"},
         inp, run_model(configs.config_diffusion, "nll", dtype=torch.float32, **kw),
         run_model(configs.config_diffusion, "nll", dtype=torch.float64, **kw))

    # 5. config_diffusion_CA (8 layers, LayerNormalization, rel-pos encoding, neighbour list as
    #    the first model layer with the chain/random criteria; torch seeded -> edges recorded)
    inp = synthetic.protein_like(36, seed=15, n_chains=2)
    inp.pop("edge_index"), inp.pop("_n_edges")
    attrs = {"CA": NODE("1x1o"), "species": NODE("1x0e"), "chain_id": NODE("1x0e"), "id": NODE("1x0e"),
             "t": ("graph", "1x0e"), "_n_nodes": ("graph", "1x0e")}
    kw = dict(inputs=inp, attrs=attrs, seed=6, out_keys=["score_CA", "node_features"], torch_seed=77)
    save("model_diffusion_CA", {"config": "config_diffusion_CA", "seed": 6, "torch_seed": 77},
         inp, run_model(configs.config_diffusion_CA, "", dtype=torch.float32, **kw),
         run_model(configs.config_diffusion_CA, "", dtype=torch.float64, **kw))


def neighbour_cases():
    """computeEdgeIndex of the genuine reference (data/compute_edge.py:38-113), fp32."""
    cases = {}
    b = synthetic.qm9_like(24, seed=21)
    cases["qm9_r5"] = (b["pos"], b["_n_nodes"], 5.0)
    cases["qm9_r4"] = (b["pos"], b["_n_nodes"], 4.0)
    p = synthetic.protein_like(300, seed=22)
    cases["protein_r8"] = (p["CA"], p["_n_nodes"], 8.0 / 25.83)
    # edge cases: single atoms, an isolated pair beyond r_max, coincident points, exact-r pairs
    pos = torch.tensor([[0.0, 0, 0], [10.0, 0, 0], [10.0, 0, 0.5], [0, 0, 0], [3.0, 4.0, 0.0], [0, 0, 0],
                        [0, 0, 0], [1.0, 1.0, 1.0], [5.0, 0.0, 0.0], [0.0, 4.9999995, 0.0]])
    cases["edge_cases_r5"] = (pos, torch.tensor([[1], [2], [2], [1], [4]]), 5.0)
    g = torch.Generator().manual_seed(23)
    cases["dense_blob"] = (torch.rand(200, 3, generator=g) * 3.0, torch.tensor([[120], [80]]), 1.25)
    arrs, meta = {}, {}
    for name, (pos, n_nodes, r) in cases.items():
        batch = Batch({"pos": ("node", "1x1o"), "_n_nodes": ("graph", "1x0e")}, pos=pos.float(), _n_nodes=n_nodes)
        d, a = computeEdgeIndex(batch.data, batch.attrs, r_max=r)
        arrs[f"{name}/pos"] = pos.float().numpy()
        arrs[f"{name}/n_nodes"] = n_nodes.numpy()
        arrs[f"{name}/edge_index"] = d["edge_index"].numpy()
        arrs[f"{name}/n_edges"] = batch.data["_n_edges"].numpy()
        meta[name] = r
        print("nbr", name, d["edge_index"].shape)
    arrs["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "neighbour_lists.npz"), **arrs)


if __name__ == "__main__":
    neighbour_cases()
    model_cases()
