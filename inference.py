#!/usr/bin/env python3
"""Inference entry point of the B200 drop-in (reference ``inference.py``: build the model of ``--config``,
load ``--model_path`` (``module.`` prefixes of DDP checkpoints stripped), run the dataset batch by batch
and dump ``--output_keys``).  Data: ``--data synthetic`` or an ``.npz`` (``pos, species, _n_nodes``);
the result is written as ``.npz`` (h5py is not available here; ``.hdf5`` paths are honoured if it is)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "equivariant-nn-zoo_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--config_spec", default="")
    ap.add_argument("--output_path", default="results.npz")
    ap.add_argument("--name", default="default")
    ap.add_argument("--model_path", default=None)
    ap.add_argument("--output_keys", default="", help="comma separated")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--data", default="synthetic")
    ap.add_argument("--n_graphs", type=int, default=512)
    ap.add_argument("--batch_size", type=int, default=None)
    flags = ap.parse_args()

    from e3_layers import configs
    from e3_layers.data import Batch, computeEdgeIndex
    from e3_layers.utils import build
    from e3b200 import synthetic

    get = getattr(configs, flags.config, None)
    assert get is not None, f"Config {flags.config} not found."
    config = get(flags.config_spec) if flags.config_spec else get()
    dev = torch.device("cuda")
    model = build(config.model_config).to(dev)
    if flags.model_path:
        state = torch.load(flags.model_path, map_location=dev)
        model.load_state_dict({(k[7:] if k.startswith("module.") else k): v for k, v in state.items()})
    model.eval()
    if flags.data == "synthetic":
        data = synthetic.qm9_like(flags.n_graphs, seed=flags.seed or 0)
    else:
        z = np.load(flags.data)
        data = {k: torch.from_numpy(z[k]) for k in ("pos", "species", "_n_nodes")}
    n = data["_n_nodes"].reshape(-1)
    starts = torch.cumsum(n, 0) - n
    bs = flags.batch_size or int(config.batch_size)
    r_max = float(config.model_config.r_max)
    attrs = {"pos": ("node", "1x1o"), "species": ("node", "1x0e"), "_n_nodes": ("graph", "1x0e")}
    keys = [k for k in flags.output_keys.split(",") if k]
    results = {k: [] for k in keys}
    for lo in range(0, n.numel(), bs):
        g = torch.arange(lo, min(lo + bs, n.numel()))
        node_idx = torch.arange(int(starts[g[0]]), int(starts[g[-1]] + n[g[-1]]))
        tensors = {"pos": data["pos"][node_idx].to(dev), "species": data["species"][node_idx].to(dev),
                   "_n_nodes": data["_n_nodes"][g].to(dev)}
        batch = Batch(dict(attrs), **tensors)
        d, a = computeEdgeIndex(batch.data, batch.attrs, r_max=r_max)
        batch.update(d)
        batch.attrs.update(a)
        out = model(Batch(batch.attrs, **batch.data))
        if not keys:
            keys = [k for k in out.data if k in ("energy", "total_energy", "forces", "dipole", "score")]
            results = {k: [] for k in keys}
        for k in keys:
            results[k].append(out[k].detach().cpu())
    arrays = {k: torch.cat(v).numpy() for k, v in results.items()}
    if flags.output_path.endswith((".hdf5", ".h5")):
        try:
            import h5py
            with h5py.File(flags.output_path, "w") as f:
                for k, v in arrays.items():
                    f.create_dataset(k, data=v)
            return
        except ImportError:
            flags.output_path = os.path.splitext(flags.output_path)[0] + ".npz"
    np.savez(flags.output_path, **arrays)
    print("wrote", flags.output_path, {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    main()
