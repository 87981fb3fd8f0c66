"""Builds and runs the PRODUCT models (e3_layers mirror on libe3b200) for the parity tests."""
import torch

import harness
from e3_layers import configs
from e3_layers.data import Batch, computeEdgeIndex
from e3_layers.utils import build
from param_init import reseed_parameters


def product_config(meta, **model_overrides):
    name, spec = meta["config"], meta.get("spec")
    if name == "config_diffusion_CA":
        spec = "no_edge_layer"
    cfg = getattr(configs, name)(spec) if spec is not None else getattr(configs, name)()
    return cfg


def build_product(meta, dtype, device):
    torch.set_default_dtype(dtype)
    try:
        model = build(product_config(meta).model_config)
        reseed_parameters(model, meta["seed"])
    finally:
        torch.set_default_dtype(torch.float32)
    return model.to(device).eval()


def run_product(model, inputs, dtype, device, pre_edge=None, edge_index=None, compute_edge=None):
    torch.set_default_dtype(dtype)
    try:
        data = harness.cast_inputs(inputs, dtype, device)
        batch = Batch(harness.attrs_for(data), **data)
        if edge_index is not None:
            batch["edge_index"] = edge_index.to(device)
        elif pre_edge is not None:
            d, a = (compute_edge or computeEdgeIndex)(batch.data, batch.attrs, **pre_edge)
            batch.update(d)
            batch.attrs.update(a)
            batch = Batch(batch.attrs, **batch.data)
        out = model(batch)
    finally:
        torch.set_default_dtype(torch.float32)
    return out
