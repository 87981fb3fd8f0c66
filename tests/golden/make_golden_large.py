"""Golden fixtures at the BASELINE workload sizes (W1 128 molecules, W2 512, W3 256,
W4 128 complete graphs, W5 1 000 residues through the neighbour-list layer), produced by the GENUINE reference
package (/root/reference/e3_layers, unmodified) on top of oracle/shims.py -- same recipe as make_golden.py.

Only fp64 outputs are kept (the truth the fp32 product is held to at 1e-5); inputs are stored too so that the
GPU box needs neither /root/reference nor the oracle for these tests.  Independent molecules are evaluated in
chunks of 16 graphs (a Batch is a concatenation without cross-graph edges, so chunking changes nothing but
the peak memory of the reference's per-edge [E, 6528] intermediates).

Run (build container only):  python tests/golden/make_golden_large.py [W1 W2 W3 W4 W5]
"""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "equivariant-nn-zoo_b200"))
sys.path.insert(0, HERE)

from e3b200 import synthetic  # noqa: E402
import make_golden as mg  # noqa: E402  (imports the reference through the shims)

configs = mg.configs
NODE = mg.NODE
ATTRS_MOL = {"pos": NODE("1x1o"), "species": NODE("1x0e"), "_n_nodes": ("graph", "1x0e")}


def _chunks(inputs, size, per_edge=()):
    """split a molecule batch into chunks of `size` graphs (node- and graph-wise tensors; per-edge tensors of
    complete graphs listed in per_edge)"""
    n = inputs["_n_nodes"].reshape(-1)
    G = n.numel()
    node_off = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(n, 0)])
    edge_off = None
    if "_n_edges" in inputs:
        edge_off = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(inputs["_n_edges"].reshape(-1), 0)])
    N = int(n.sum())
    for g0 in range(0, G, size):
        g1 = min(G, g0 + size)
        a, b = int(node_off[g0]), int(node_off[g1])
        out = {}
        for k, v in inputs.items():
            if k == "edge_index":
                e0, e1 = int(edge_off[g0]), int(edge_off[g1])
                out[k] = v[:, e0:e1] - a
            elif k in per_edge:
                e0, e1 = int(edge_off[g0]), int(edge_off[g1])
                out[k] = v[e0:e1]
            elif v.shape[0] == N and k not in ("_n_nodes", "_n_edges", "t"):
                out[k] = v[a:b]
            else:
                out[k] = v[g0:g1]
        yield out


def _sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).hexdigest()


def run_chunked(cfg_fn, spec, inputs, attrs, seed, out_keys, graph_keys, pre_edge=None, chunk=16, per_edge=()):
    outs = {k: [] for k in out_keys}
    eis, off = [], 0
    for part in _chunks(inputs, chunk, per_edge):
        r = mg.run_model(cfg_fn, spec, inputs=part, attrs=attrs, seed=seed, dtype=torch.float64, out_keys=out_keys,
                         pre_edge=pre_edge)
        for k in out_keys:
            outs[k].append(torch.from_numpy(r[k]))
        eis.append(torch.from_numpy(r["edge_index"]) + off)
        off += part["pos"].shape[0]
    res = {k: torch.cat(v) for k, v in outs.items()}
    res["edge_index"] = torch.cat(eis, dim=1)
    return res


def save(name, meta, inputs, out64):
    arrs = {"in/" + k: v.numpy() for k, v in inputs.items()}
    ei = out64.pop("edge_index")
    meta = dict(meta, n_edges=int(ei.shape[1]), edge_index_sha256=_sha(ei.long()))
    for k, v in out64.items():
        arrs["out64/" + k] = v.numpy()
    arrs["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(name, meta, {k: v.shape for k, v in arrs.items() if k != "meta"}, os.path.getsize(path), "bytes", flush=True)


def w1():
    inp = synthetic.qm9_like(128, seed=101)
    out = run_chunked(configs.config_energy_force, None, inp, ATTRS_MOL, 1, ["energy", "forces"], ["energy"],
                      pre_edge={"r_max": 5.0})
    save("large_W1_energy_force_128", {"config": "config_energy_force", "seed": 1, "r_max": 5.0}, inp, out)


def w2():
    inp = synthetic.qm9_like(512, seed=0)      # the very batch bench.py times
    out = run_chunked(configs.config_energy_force, None, inp, ATTRS_MOL, 1, ["energy", "forces"], ["energy"],
                      pre_edge={"r_max": 5.0})
    save("large_W2_energy_force_512", {"config": "config_energy_force", "seed": 1, "r_max": 5.0}, inp, out)


def w3():
    inp = synthetic.qm9_like(256, seed=103, species_choices=tuple(range(1, 18)))
    out = run_chunked(configs.config_dipole, None, inp, ATTRS_MOL, 3, ["dipole"], [], pre_edge={"r_max": 5.0})
    save("large_W3_dipole_256", {"config": "config_dipole", "seed": 3, "r_max": 5.0}, inp, out)


def w4():
    inp = synthetic.diffusion_like(128, seed=104)
    attrs = dict(ATTRS_MOL, t=("graph", "1x0e"), bond_type=("edge", "1x0e"), _n_edges=("graph", "1x0e"))
    out = run_chunked(configs.config_diffusion, "", inp, attrs, 4, ["score"], [], per_edge=("bond_type",))
    save("large_W4_diffusion_128", {"config": "config_diffusion", "seed": 4, "spec": ""}, inp, out)


def w5():
    n_res, torch_seed = 1000, 77
    inp = synthetic.protein_like(n_res, seed=105, n_chains=4)
    inp.pop("edge_index"), inp.pop("_n_edges")
    attrs = {"CA": NODE("1x1o"), "species": NODE("1x0e"), "chain_id": NODE("1x0e"), "id": NODE("1x0e"),
             "t": ("graph", "1x0e"), "_n_nodes": ("graph", "1x0e")}
    # the reference's criteria draws torch.rand(n^2) from the CPU generator (config_diffusion_CA.py:58-64); seeded
    # right before the forward by run_model, so the test side reproduces the same uniforms from the same seed
    with torch.no_grad():
        r = mg.run_model(configs.config_diffusion_CA, "", inputs=inp, attrs=attrs, seed=6, dtype=torch.float64,
                         out_keys=["score_CA"], torch_seed=torch_seed)
    out = {k: torch.from_numpy(v) for k, v in r.items()}
    save("large_W5_diffusion_CA_1000", {"config": "config_diffusion_CA", "seed": 6, "torch_seed": torch_seed,
                                        "p_random": 0.02, "n_res": n_res}, inp, out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["W1", "W3", "W4", "W5", "W2"]
    for w in which:
        t0 = time.time()
        {"W1": w1, "W2": w2, "W3": w3, "W4": w4, "W5": w5}[w]()
        print(w, "done in %.0f s" % (time.time() - t0), flush=True)
