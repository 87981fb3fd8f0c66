"""``CondensedDataset``: a ``Batch`` holding a whole dataset, with loading, key mapping, per-item preprocessing and
statistics (API of the reference's ``e3_layers/data/dataset.py:22-330``).

Differences that matter on the B200 path:

* the tensors stay CONCATENATED (one tensor per key for the whole dataset, "pre-concatenated shards"); batches are cut
  out of them by ``e3_layers.data.dataloader.DevicePipeline`` with vectorised gathers instead of per-molecule Python
  objects collated by DataLoader workers;
* a preprocess function follows the LAYER contract ``func(data, attrs) -> (new tensors, attrs)`` and its result is
  MERGED into the item.  The reference replaces the item's tensors with the function's return value
  (``dataset.py:115-117``), which for ``computeEdgeIndex`` -- which returns only ``{"edge_index": ...}``
  (``compute_edge.py:110-113``) -- drops the positions and species; merged, the same function works both as a model layer
  and as a dataset preprocess;
* files: ``.npz`` (always) and HDF5 (when ``h5py`` is importable; same key / attribute conventions as the reference)."""
import logging
import os
import re
from inspect import signature

import numpy as np
import torch

from ..utils import keyMap
from .batch import Batch

# atomic symbols in order of atomic number (index = Z); the reference takes them from ase.atom.atomic_numbers
_SYMBOLS = ("X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y Zr Nb "
            "Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg "
            "Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md No Lr").split()


def _normalise(item):
    item = torch.as_tensor(item)
    if item.dtype == torch.int32:
        return item.long()
    if item.dtype == torch.float64:
        return item.float()
    return item


def _load_file(path):
    logging.info("Loading %s", path)
    data, attrs = {}, {}
    if path.endswith(".npz"):
        with np.load(path, allow_pickle=True) as z:
            for key in z.files:
                if key.startswith("__attrs__/"):
                    v = z[key].tolist()
                    attrs[key[len("__attrs__/"):]] = (v[0], v[1])
                else:
                    data[key] = _normalise(z[key])
    else:
        import h5py  # optional: only for the reference's own .hdf5 files

        with h5py.File(path, "r") as f:
            for key in f.keys():
                data[key] = _normalise(f[key][:])
            for key in f.attrs.keys():
                attrs[key] = f.attrs[key]
    return data, attrs


def save_npz(path, batch):
    """writes a Batch / CondensedDataset as the .npz this module loads (tensors + ``__attrs__/key`` entries)"""
    arrs = {k: v.detach().cpu().numpy() for k, v in batch.data.items() if k not in ("_node_segment", "_edge_segment")}
    for k, v in batch.attrs.items():
        if k not in ("_node_segment", "_edge_segment"):
            arrs["__attrs__/" + k] = np.array([v[0], str(v[1])], dtype=object)
    np.savez(path, **arrs)


class CondensedDataset(Batch):
    def __init__(self, path=None, data=None, attrs=None, key_map=None, type_names=None, preprocess=None, **kwargs):
        """path: a file, a list of files, a directory, or ``directory:regular_expression`` (reference ``dataset.py:27-45``)"""
        data, attrs = ({} if data is None else data), ({} if attrs is None else attrs)
        if path is not None:
            data, attrs = CondensedDataset.load(path)
            if isinstance(data, list):
                data = Batch.from_data_list(data, attrs).data
        attrs = {k: (v[0], v[1]) for k, v in keyMap(dict(attrs), key_map or {}).items()}
        super().__init__(attrs, **keyMap(dict(data), key_map or {}))
        self.type_names = list(_SYMBOLS if type_names is None else type_names)
        self.preprocess = list(preprocess or [])
        self.kwargs = kwargs

    @staticmethod
    def load(path):
        if isinstance(path, str):
            parts = path.split(":")
            path, regexp = (parts[0], re.compile(parts[1])) if len(parts) == 2 else (parts[0], None)
            if os.path.isdir(path):
                data, attrs = [], {}
                for root, _, files in sorted(os.walk(path)):
                    for name in sorted(files):
                        file = os.path.join(root, name)
                        if regexp is not None and regexp.match(file) is None:
                            continue
                        d, a = _load_file(file)
                        data.append(d)
                        attrs.update(a)
            else:
                data, attrs = _load_file(path)
        else:
            data, attrs = [], {}
            for item in path:
                d, a = CondensedDataset.load(item)
                data += d if isinstance(d, list) else [d]
                attrs.update(a)
        if isinstance(data, list) and not data:
            logging.warning("No dataset file is found in %s.", path)
        return data, attrs

    @staticmethod
    def apply_preprocess(item, funcs):
        """runs the preprocess chain on a Data / Batch.  One-argument functions map the item; two-argument functions
        follow the layer contract and their outputs are merged into the item."""
        for func in funcs:
            if len(signature(func).parameters) == 1:
                item = func(item)
            else:
                tensors, attrs = func(item.data, item.attrs)
                item.attrs.update(attrs)
                item.update(tensors)
        return item

    def __getitem__(self, idx):
        if isinstance(idx, str):
            return self.data[idx]
        if isinstance(idx, (int, np.integer)):
            return CondensedDataset.apply_preprocess(self.get(int(idx)).clone(), self.preprocess)
        return self.index_select(idx)

    def index_select(self, idx):
        batch = super().index_select(idx)
        return CondensedDataset(type_names=self.type_names, preprocess=self.preprocess, data=batch.data, attrs=batch.attrs)

    # -- statistics (reference dataset.py:139-330), on the concatenated tensors -------------------------------------
    def statistics(self, fields, stride=1, unbiased=True):
        ds = self if stride == 1 else self.index_select(list(range(0, len(self), stride)))
        seg = ds.nodeSegment() if "_n_nodes" in ds.data else None
        out = []
        for field in fields:
            key = field.split("-")[0]
            mode = field[len(key) + 1:]
            arr = ds[key]
            arr = arr if arr.is_floating_point() or mode == "count" else arr.float()
            is_per = ds.attrs[key][0]
            if mode == "count":
                out.append(torch.unique(arr.flatten(), return_counts=True, sorted=True))
            elif mode == "rms":
                out.append((torch.sqrt(torch.mean(arr * arr)),))
            elif mode == "mean_std":
                out.append((arr.mean(dim=0), arr.std(dim=0, unbiased=unbiased)))
            elif mode.startswith("per-node-"):
                if is_per != "graph":
                    raise ValueError(f"It doesn't make sense to ask for `{mode}` since `{field}` is not per-graph")
                per = arr / ds["_n_nodes"].reshape(-1, 1).to(arr.dtype)
                sub = mode[len("per-node-"):]
                if sub == "mean_std":
                    out.append((per.mean(dim=0), per.std(dim=0, unbiased=unbiased)))
                elif sub == "rms":
                    out.append((torch.sqrt(torch.mean(per * per)),))
                else:
                    raise NotImplementedError(f"Cannot handle statistics mode {mode}")
            elif mode.startswith("per-"):
                _, tkey, sub = mode.split("-")
                out.append(self._per_species(sub, arr, is_per, seg, ds[tkey].reshape(-1), unbiased))
            else:
                raise NotImplementedError(f"Cannot handle statistics mode {mode}")
        return out

    def _per_species(self, mode, arr, is_per, seg, types, unbiased):
        n_types = len(self.type_names)
        if is_per == "node":
            if mode == "rms":
                sq = torch.zeros(n_types, *arr.shape[1:]).index_add_(0, types, arr * arr)
                cnt = torch.bincount(types, minlength=n_types).clamp_min(1).view(-1, *([1] * (arr.dim() - 1)))
                return (torch.sqrt(sq / cnt),)
            if mode == "mean_std":
                cnt = torch.bincount(types, minlength=n_types).view(-1, *([1] * (arr.dim() - 1)))
                mean = torch.zeros(n_types, *arr.shape[1:]).index_add_(0, types, arr) / cnt.clamp_min(1)
                var = torch.zeros(n_types, *arr.shape[1:]).index_add_(0, types, (arr - mean[types]) ** 2)
                return mean, torch.sqrt(var / (cnt - (1 if unbiased else 0)).clamp_min(1))
            raise NotImplementedError(f"Statistics mode {mode} isn't yet implemented for per_species_mean_std")
        if is_per == "graph" and mode == "mean_std":
            # least squares of the per-graph value on the composition (reference: utils.solver on the bincount matrix)
            G = int(seg.max()) + 1 if seg.numel() else 0
            comp = torch.zeros(G, n_types).index_put_((seg, types), torch.ones(types.numel()), accumulate=True)
            used = comp.sum(0) > 0
            sol = torch.linalg.lstsq(comp[:, used].double(), arr.double()).solution
            mean = torch.zeros(n_types, *arr.shape[1:], dtype=torch.float64)
            mean[used] = sol
            res = arr.double() - comp.double() @ mean
            std = res.std(dim=0, unbiased=unbiased)
            return mean.float(), std.float()
        raise NotImplementedError(f"per-species statistics of a per-{is_per} quantity in mode {mode}")
