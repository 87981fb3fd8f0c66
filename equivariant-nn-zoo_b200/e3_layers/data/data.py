"""``Data``: a dict of tensors plus ``attrs[key] = (is_per, irreps)`` describing whether a tensor
is per node / edge / graph and its irreps (API of the reference's ``e3_layers/data/data.py:13-238``)."""
import copy
import re

import torch

from e3b200.irreps import Irreps


def irreps_dim(irreps):
    if isinstance(irreps, int):
        return irreps
    if isinstance(irreps, str) and irreps.isdigit():
        return int(irreps)
    return Irreps(irreps).dim


class Data(object):
    def __init__(self, attrs=None, **tensors):
        self.attrs = {} if attrs is None else attrs
        self.data = {}
        self.device = None
        for key, value in tensors.items():
            self._store(key, value)
        self.computeSums()

    # -- bookkeeping --------------------------------------------------------------------------
    def __cat_dim__(self, key):
        return -1 if re.search("(index|face)", key) else 0

    def num_dims(self, key):
        if key in self.attrs:
            return irreps_dim(self.attrs[key][1])
        return None

    def computeSums(self):
        for key, tensor in self.data.items():
            kind = self.attrs.get(key, (None,))[0]
            n = tensor.shape[self.__cat_dim__(key)] if tensor.dim() else 0
            if kind == "node":
                self.n_nodes = n
            elif kind == "edge":
                self.n_edges = n
            elif kind == "graph":
                self.n_graphs = n

    def _store(self, key, item):
        if not isinstance(item, torch.Tensor):
            item = torch.tensor(item)
        dim = self.num_dims(key)
        if dim is not None and not (item.dim() == 2 and item.shape[-1] == dim):
            item = item.reshape(-1, dim)
        self.data[key] = item
        if item.is_cuda and self.device is None:
            self.device = item.device

    # -- mapping protocol ---------------------------------------------------------------------
    def __getitem__(self, key):
        return self.data[key]

    def __setitem__(self, key, item):
        self._store(key, item)
        self.computeSums()

    def __contains__(self, key):
        return key in self.data

    def __len__(self):
        return len(self.data)

    def keys(self):
        return self.data.keys()

    def items(self):
        return list(self.data.items())

    def update(self, other):
        for key, value in other.items():
            self._store(key, value)
        self.computeSums()

    def pop(self, key):
        self.data.pop(key, None)
        self.attrs.pop(key, None)

    def __call__(self, *keys):
        for key in (keys or sorted(self.keys())):
            if key in self:
                yield key, self[key]

    @property
    def num_edges(self):
        return self["edge_index"].shape[-1]

    # -- tensor movement ----------------------------------------------------------------------
    def apply(self, func, *keys):
        for key in keys:
            v = self.data[key]
            self.data[key] = func(v) if torch.is_tensor(v) else v
        return self

    def to(self, device, **kwargs):
        self.device = device
        return self.apply(lambda t: t.to(device, **kwargs), *self.keys())

    def cpu(self):
        return self.to("cpu")

    def cuda(self, device=None, non_blocking=False):
        self.device = torch.device("cuda" if device is None else device)
        return self.apply(lambda t: t.cuda(device=device, non_blocking=non_blocking), *self.keys())

    def contiguous(self):
        return self.apply(lambda t: t.contiguous(), *self.keys())

    def pin_memory(self):
        return self.apply(lambda t: t.pin_memory(), *self.keys())

    def clone(self):
        new = self.__class__(copy.deepcopy(self.attrs),
                             **{k: (v.clone() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in self.data.items()})
        new.device = self.device
        return new

    def __repr__(self):
        shapes = {k: (tuple(v.shape), v.dtype) for k, v in self.data.items()}
        return f"attrs:{self.attrs}\n tensors:{shapes}"

    def dumpHDF5(self, path):
        import h5py  # optional dependency, only for dataset export

        with h5py.File(path, "w") as f:
            for key in self.keys():
                if key not in ("_node_segment", "_edge_segment"):
                    f[key] = self[key].to("cpu")
            for key, value in self.attrs.items():
                if key not in ("_node_segment", "_edge_segment"):
                    f.attrs[key] = value
