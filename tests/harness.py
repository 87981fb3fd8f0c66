"""Shared helpers for the parity tests: golden loading, oracle model construction."""
import json
import os

import numpy as np
import torch

from oracle import ref_configs, ref_layers
from param_init import reseed_parameters

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

NODE_ATTRS = {
    "pos": ("node", "1x1o"), "species": ("node", "1x0e"), "_n_nodes": ("graph", "1x0e"),
    "t": ("graph", "1x0e"), "bond_type": ("edge", "1x0e"), "_n_edges": ("graph", "1x0e"),
    "CA": ("node", "1x1o"), "chain_id": ("node", "1x0e"), "id": ("node", "1x0e"),
}


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    out = {"meta": meta, "in": {}, "out32": {}, "out64": {}}
    for k in z.files:
        if k == "meta":
            continue
        grp, key = k.split("/", 1)
        out[grp][key] = torch.from_numpy(z[k])
    return out


def oracle_config(meta, **overrides):
    name = meta["config"]
    if name == "config_diffusion":
        return ref_configs.config_diffusion(nll=(meta.get("spec") == "nll"), **overrides)
    return getattr(ref_configs, name)(**overrides)


def build_oracle(meta, dtype, **overrides):
    torch.set_default_dtype(dtype)
    try:
        model = ref_layers.build(oracle_config(meta, **overrides))
        reseed_parameters(model, meta["seed"])
    finally:
        torch.set_default_dtype(torch.float32)
    return model.eval()


def cast_inputs(inputs, dtype, device="cpu"):
    return {k: (v.to(dtype) if v.is_floating_point() else v.clone()).to(device) for k, v in inputs.items()}


def attrs_for(inputs):
    return {k: NODE_ATTRS[k] for k in inputs if k in NODE_ATTRS}


def run_oracle(model, inputs, dtype, pre_edge=None, edge_index=None):
    """Runs an oracle model (OracleNetwork or GradientOutput) the way make_golden ran the
    reference: optional dataset-style neighbour list first, then the model."""
    torch.set_default_dtype(dtype)
    try:
        data = cast_inputs(inputs, dtype)
        attrs = attrs_for(data)
        if edge_index is not None:
            data["edge_index"] = edge_index
        elif pre_edge is not None:
            n = data["_n_nodes"].reshape(-1)
            data["_node_segment"] = torch.repeat_interleave(torch.arange(len(n)), n)
            d, attrs = ref_layers.computeEdgeIndex(data, attrs, **pre_edge)
            data.update(d)
        out, _ = model(data, attrs)
    finally:
        torch.set_default_dtype(torch.float32)
    return out


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a.detach() - b.detach()).abs().max() / b.detach().abs().max().clamp_min(1e-300))
