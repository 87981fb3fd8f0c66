"""``Batch``: graphs concatenated without padding (reference ``e3_layers/data/batch.py:10-201``).
Segment vectors (graph id of every node / edge) are built on the tensors' own device with
``repeat_interleave`` instead of Python loops + CPU tensors."""
from collections.abc import Sequence

import numpy as np
import torch

from .data import Data, irreps_dim


def _segments(counts, total=None):
    """graph id of every node / edge.  With `total` (the number of rows, known from a per-node / per-edge
    tensor) repeat_interleave does not have to read the counts back: no host synchronisation."""
    counts = counts.reshape(-1)
    ids = torch.arange(counts.numel(), device=counts.device)
    if total is None:
        return torch.repeat_interleave(ids, counts)
    return torch.repeat_interleave(ids, counts, output_size=total)


def _tag(counts):
    return (counts.data_ptr(), counts._version, counts.numel())


class Batch(Data):
    def __init__(self, attrs=None, **tensors):
        super().__init__(attrs, **tensors)
        if "_n_nodes" in self.data:
            self.nodeSegment()
        if "_n_edges" in self.data:
            self.edgeSegment()

    def _rows(self, kind):
        """number of nodes / edges if some tensor of that kind tells it, else None"""
        if kind == "edge" and "edge_index" in self.data:
            return int(self.data["edge_index"].shape[1])
        for key, (per, _) in self.attrs.items():
            if per == kind and key in self.data and not key.startswith("_"):
                return int(self.data[key].shape[0])
        return None

    def _segment(self, kind, counts_key, seg_key):
        counts = self.data[counts_key]
        total = self._rows(kind)
        old = self.data.get(seg_key)
        # a Batch rebuilt from another Batch's tensors keeps the segment vector computed for the same counts
        if old is not None and getattr(old, "_e3b_counts", None) == _tag(counts) and (total is None or old.shape[0] == total):
            return old
        seg = _segments(counts, total)
        seg._e3b_counts = _tag(counts)
        self.data[seg_key] = seg
        return seg

    def nodeSegment(self):
        return self._segment("node", "_n_nodes", "_node_segment")

    def edgeSegment(self):
        return self._segment("edge", "_n_edges", "_edge_segment")

    def computeCumsums(self):
        for what in ("node", "edge"):
            key = f"_n_{what}s"
            if key in self.data and not hasattr(self, f"{what}_cumsum"):
                counts = self.data[key].reshape(-1).cpu()
                cs = torch.zeros(counts.numel() + 1, dtype=torch.long)
                cs[1:] = torch.cumsum(counts, 0)
                setattr(self, f"{what}_cumsum", cs)
                self.n_graphs = counts.numel()
                setattr(self, f"n_{what}s", int(cs[-1]))

    @classmethod
    def from_data_list(cls, lst, attrs=None):
        attrs = {} if attrs is None else attrs
        attrs["_n_nodes"] = ("graph", "1x0e")
        attrs["_n_edges"] = ("graph", "1x0e")
        node_key = next((k for k in lst[0].keys() if k in attrs and attrs[k][0] == "node"), None)
        items = []
        for item in lst:
            item = dict(item.items()) if not isinstance(item, dict) else dict(item)
            if "_n_nodes" not in item:
                assert node_key is not None, "Unable to infer the amount of nodes."
                item["_n_nodes"] = torch.full((1, 1), torch.as_tensor(item[node_key]).shape[0], dtype=torch.long)
            item["_n_nodes"] = torch.as_tensor(item["_n_nodes"]).view(-1, 1)
            if "_n_edges" not in item and "edge_index" in item:
                item["_n_edges"] = torch.full((1, 1), torch.as_tensor(item["edge_index"]).shape[1], dtype=torch.long)
            elif "_n_edges" in item:
                item["_n_edges"] = torch.as_tensor(item["_n_edges"]).view(-1, 1)
            items.append(item)
        out = {"_n_nodes": torch.cat([it["_n_nodes"] for it in items])}
        if "_n_edges" in items[0]:
            out["_n_edges"] = torch.cat([it["_n_edges"] for it in items])
        for key in items[0].keys():
            if key in out or key in ("_node_segment", "_edge_segment"):
                continue
            parts = [torch.as_tensor(it[key]) for it in items]
            if key == "edge_index":
                offs = torch.cumsum(torch.tensor([0] + [int(it["_n_nodes"].sum()) for it in items[:-1]]), 0)
                out[key] = torch.cat([p.long() + o for p, o in zip(parts, offs)], dim=-1)
                continue
            if key in attrs:
                dim = irreps_dim(attrs[key][1])
                parts = [p.reshape(-1, dim) for p in parts]
            cat = torch.cat(parts, dim=0)
            out[key] = cat.long() if cat.dtype in (torch.int64, torch.int32, torch.int16, torch.int8) else cat.float()
        return cls(attrs, **out)

    def get(self, idx):
        self.computeCumsums()
        dic = {}
        for key, value in self.data.items():
            if key == "edge_index":
                a, b = self.edge_cumsum[idx], self.edge_cumsum[idx + 1]
                dic[key] = value[:, a:b] - self.node_cumsum[idx]
            if key not in self.attrs:
                continue
            kind = self.attrs[key][0]
            if kind == "graph":
                a, b = idx, idx + 1
            else:
                cs = self.node_cumsum if kind == "node" else self.edge_cumsum
                a, b = cs[idx], cs[idx + 1]
            dic[key] = value[a:b]
        return Data(self.attrs, **dic)

    def index_select(self, idx):
        if isinstance(idx, slice):
            idx = list(range(self.n_graphs)[idx])
        elif isinstance(idx, torch.Tensor):
            idx = (idx.flatten().nonzero().flatten() if idx.dtype == torch.bool else idx.flatten()).tolist()
        elif isinstance(idx, np.ndarray):
            idx = (idx.flatten().nonzero()[0] if idx.dtype == bool else idx.flatten()).tolist()
        elif not (isinstance(idx, Sequence) and not isinstance(idx, str)):
            raise IndexError(f"invalid batch index of type {type(idx).__name__}")
        picked = [self.get(i) for i in idx]
        return Batch.from_data_list(picked, picked[0].attrs)

    def __getitem__(self, idx):
        if isinstance(idx, str):
            return self.data[idx]
        if isinstance(idx, (int, np.integer)):
            return self.get(idx)
        return self.index_select(idx)

    def __setitem__(self, key, item):
        if isinstance(key, int):
            raise NotImplementedError("Setting item with an integer index is not supported for Batch.")
        super().__setitem__(key, item)

    def __len__(self):
        if "_n_nodes" in self.data:
            return self.data["_n_nodes"].shape[0]
        return getattr(self, "n_graphs", 0)
