"""Shared pieces of the sampler / score-matching parity tests: recorded noise fed identically to the
oracle (oracle/ref_sde.py) and to the product (e3_layers.run)."""
import torch

import harness
from e3_layers.data import Batch
from oracle import ref_sde


class Noise:
    """pre-drawn standard normal tensors of one shape, replayed in order"""

    def __init__(self, shape, count, seed):
        g = torch.Generator().manual_seed(seed)
        self.draws = [torch.randn(shape, generator=g, dtype=torch.float64) for _ in range(count)]
        self.i = 0

    def __iter__(self):
        return self

    def __next__(self):
        z = self.draws[self.i]
        self.i += 1
        return z

    def randn_like(self, x):
        return next(self).to(x)


def oracle_model_fn(meta, inputs):
    """model_fn for ref_sde: evaluates the fp64 oracle network on the diffused key / t of `data`"""
    model = harness.build_oracle(meta, torch.float64)

    def fn(data):
        cur = dict(inputs)
        for k in ("pos", "CA", "t"):
            if k in data:
                cur[k] = data[k]
        ei = cur.pop("edge_index")
        return harness.run_oracle(model, cur, torch.float64, edge_index=ei)

    return fn


def oracle_data(inputs, key="pos"):
    n = inputs["_n_nodes"].reshape(-1)
    return {key: inputs[key].double(), "t": inputs["t"].double(), "_n_nodes": inputs["_n_nodes"],
            "_node_segment": torch.repeat_interleave(torch.arange(len(n)), n)}


def product_batch(inputs, dtype, device):
    data = harness.cast_inputs(inputs, dtype, device)
    return Batch(harness.attrs_for(data), **data)


__all__ = ["Noise", "oracle_model_fn", "oracle_data", "product_batch", "ref_sde"]
