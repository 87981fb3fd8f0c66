"""Dataset -> device pipeline (reference ``e3_layers/data/dataloader.py:13-119``: a torch ``DataLoader`` whose worker
processes run the per-molecule preprocess -- the neighbour list -- on the CPU and a ``Collater`` that concatenates
Python ``Data`` objects).

B200 design: the dataset stays ONE concatenated tensor per key.  ``DevicePipeline`` cuts a batch out of it with
vectorised gathers (index arithmetic on the host from the graph sizes, which the host owns anyway; no Python loop over
molecules, no worker processes):

* ``resident=True``  -- the whole dataset lives in HBM (QM9: ~50 MB of 180 GB); a batch is a handful of device gathers;
* ``resident=False`` -- the dataset lives in pinned host memory; rows are gathered into pinned staging buffers and
  copied by the copy engine on a side stream, ``prefetch`` batches ahead of the consumer (events order the hand-over).

The preprocess chain (e.g. ``computeEdgeIndex``: the radius-graph kernel) then runs ON THE DEVICE on the whole batch, with
the layer contract honoured -- its outputs are merged into the batch (``CondensedDataset.apply_preprocess``).
``Collater`` / ``DataLoader`` / ``getDataIters`` keep the reference's names and semantics (rank split of a path list,
train / validation split, seeded shuffling, ``drop_last``, endless iterators)."""
import math

import numpy as np
import torch

from .batch import Batch
from .dataset import CondensedDataset


class Collater(object):
    """concatenates a list of ``Data`` (reference ``dataloader.py:13-30``); the pipeline does not need it, callers that
    build batches from individual items do"""

    @classmethod
    def for_dataset(cls, dataset):
        return cls()

    def collate(self, batch):
        return Batch.from_data_list(batch, attrs=batch[0].attrs)

    def __call__(self, batch):
        return self.collate(batch)


def _ranges(starts, counts):
    """concatenation of arange(starts[i], starts[i] + counts[i]) without a Python loop (numpy int64)"""
    total = int(counts.sum())
    if total == 0:
        return np.zeros(0, dtype=np.int64)
    out_start = np.cumsum(counts) - counts
    return np.repeat(starts - out_start, counts) + np.arange(total, dtype=np.int64)


class DevicePipeline:
    def __init__(self, dataset, batch_size=1, shuffle=False, drop_last=False, generator=None, device=None, resident=None,
                 preprocess=None, rank=0, world_size=1, prefetch=2, resident_limit_bytes=8 << 30, indices=None):
        """dataset: CondensedDataset / Batch on the host; ``indices``: the graphs to draw from (default: all).  With
        world_size > 1 every rank iterates the same global order (same generator seed) and takes graphs rank,
        rank + world, ... of each global batch of ``batch_size * world_size`` graphs."""
        self.attrs = dict(dataset.attrs)
        self.preprocess = list(dataset.preprocess if preprocess is None and hasattr(dataset, "preprocess") else (preprocess or []))
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), bool(shuffle), bool(drop_last)
        self.generator = generator if generator is not None else torch.Generator()
        self.rank, self.world = int(rank), int(world_size)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.prefetch = max(1, int(prefetch))
        skip = ("_node_segment", "_edge_segment")
        host = {k: v for k, v in dataset.data.items() if k not in skip}
        self.total_graphs = int(host["_n_nodes"].shape[0])
        self.subset = np.arange(self.total_graphs) if indices is None else np.asarray(torch.as_tensor(indices).reshape(-1).tolist(), dtype=np.int64)
        self.n_graphs = int(self.subset.size)
        self.n_nodes = host["_n_nodes"].reshape(-1).numpy().astype(np.int64)
        self.node_start = np.cumsum(self.n_nodes) - self.n_nodes
        self.has_edges = "edge_index" in host
        if self.has_edges:
            self.n_edges = host["_n_edges"].reshape(-1).numpy().astype(np.int64)
            self.edge_start = np.cumsum(self.n_edges) - self.n_edges
        self.kind = {}
        for k in host:
            if k == "edge_index":
                self.kind[k] = "edge_index"
            elif k in self.attrs:
                self.kind[k] = self.attrs[k][0]
            else:
                self.kind[k] = "graph" if host[k].shape[0] == self.total_graphs else "node"
        nbytes = sum(v.numel() * v.element_size() for v in host.values())
        self.resident = (nbytes <= resident_limit_bytes) if resident is None else bool(resident)
        on_gpu = self.device.type == "cuda"
        if self.resident:
            self.store = {k: v.to(self.device) for k, v in host.items()}
        else:
            self.store = {k: (v.contiguous().pin_memory() if on_gpu else v.contiguous()) for k, v in host.items()}
            self.copy_stream = torch.cuda.Stream(self.device) if on_gpu else None
            self.staging = [None] * (self.prefetch + 1)            # ring of pinned staging dicts, grown on demand
        self.bytes = nbytes

    def __len__(self):
        per = self.batch_size * self.world
        return self.n_graphs // per if self.drop_last else math.ceil(self.n_graphs / per)

    # -- which graphs --------------------------------------------------------------------------------------------
    def _order(self):
        if self.shuffle:
            return self.subset[torch.randperm(self.n_graphs, generator=self.generator).numpy()]
        return self.subset

    def batches_of_indices(self):
        """graph ids of this rank's batches of one epoch (ascending within a batch: dataset order, like the
        reference's collation of a sorted index list)"""
        order = self._order()
        per = self.batch_size * self.world
        for b in range(len(self)):
            g = np.sort(order[b * per:(b + 1) * per])
            yield g[self.rank::self.world]

    def endless(self, skip=0):
        """batches for ever, epoch after epoch (reference ``autoReset``); ``skip`` batches are drawn but not built
        (resuming a run replays the data order without touching the data)"""
        if len(self) == 0:
            raise ValueError(f"no batch can be formed: {self.n_graphs} graphs, batch_size {self.batch_size} x {self.world} "
                             f"rank(s), drop_last={self.drop_last}")
        while True:
            todo = list(self.batches_of_indices())
            if skip >= len(todo):
                skip -= len(todo)
                continue
            todo, skip = todo[skip:], 0
            yield from self._run(todo)

    # -- index arithmetic (host) --------------------------------------------------------------------------------
    def _plan(self, graphs):
        graphs = np.asarray(graphs, dtype=np.int64)
        cn = self.n_nodes[graphs]
        plan = {"graph": graphs, "node": _ranges(self.node_start[graphs], cn)}
        if self.has_edges:
            ce = self.n_edges[graphs]
            plan["edge"] = _ranges(self.edge_start[graphs], ce)
            # edge_index is stored with dataset-global node ids: shift every graph's block to its place in the batch
            new_node_start = np.cumsum(cn) - cn
            plan["edge_shift"] = np.repeat(new_node_start - self.node_start[graphs], ce)
        return plan

    # -- gathers -------------------------------------------------------------------------------------------------
    def _gather_resident(self, plan):
        dev = self.device
        idx = {k: torch.from_numpy(v).to(dev, non_blocking=True) for k, v in plan.items()}
        out = {}
        for k, v in self.store.items():
            kind = self.kind[k]
            if kind == "edge_index":
                out[k] = v.index_select(1, idx["edge"]) + idx["edge_shift"]
            else:
                out[k] = v.index_select(0, idx[kind])
        return out, None

    def _gather_pinned(self, plan, slot):
        idx = {k: torch.from_numpy(v) for k, v in plan.items()}
        stage = self.staging[slot]
        if stage is None:
            stage = self.staging[slot] = {"event": None, "buf": {}}
        if stage["event"] is not None:
            stage["event"].synchronize()                     # the copy that last used these buffers has finished
        host = {}
        for k, v in self.store.items():
            kind = self.kind[k]
            ix = idx["edge" if kind == "edge_index" else kind]
            dim = 1 if kind == "edge_index" else 0
            shape = list(v.shape)
            shape[dim] = ix.numel()
            buf = stage["buf"].get(k)
            if buf is None or buf.numel() < int(np.prod(shape)):
                n = max(int(np.prod(shape)) * 5 // 4, 1)
                buf = torch.empty(n, dtype=v.dtype)
                buf = buf.pin_memory() if self.copy_stream is not None else buf
                stage["buf"][k] = buf
            dst = buf[:int(np.prod(shape))].view(shape)
            torch.index_select(v, dim, ix, out=dst)
            if kind == "edge_index":
                dst += idx["edge_shift"]
            host[k] = dst
        if self.copy_stream is None:
            return {k: v.clone() for k, v in host.items()}, None
        with torch.cuda.stream(self.copy_stream):
            out = {k: v.to(self.device, non_blocking=True) for k, v in host.items()}
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        stage["event"] = ev
        return out, ev

    def gather(self, graphs, slot=0):
        """the batch of the given graph ids as device tensors (+ the event that orders the copies, if any)"""
        plan = self._plan(graphs)
        return self._gather_resident(plan) if self.resident else self._gather_pinned(plan, slot)

    def make_batch(self, tensors, event=None):
        if event is not None:
            torch.cuda.current_stream(self.device).wait_event(event)
            for t in tensors.values():                       # allocated on the copy stream, consumed on the current one
                t.record_stream(torch.cuda.current_stream(self.device))
        batch = Batch(dict(self.attrs), **tensors)
        return CondensedDataset.apply_preprocess(batch, self.preprocess)

    def __iter__(self):
        return self._run(list(self.batches_of_indices()))

    def _run(self, todo):
        queue = []
        nxt = 0
        while nxt < len(todo) or queue:
            while nxt < len(todo) and len(queue) < (1 if self.resident else self.prefetch):
                queue.append(self.gather(todo[nxt], slot=nxt % (self.prefetch + 1)))
                nxt += 1
            tensors, ev = queue.pop(0)
            yield self.make_batch(tensors, ev)


class DataLoader(DevicePipeline):
    """the reference's name for the thing one iterates over (``dataloader.py:32-49``); accepts and ignores the worker
    arguments of ``torch.utils.data.DataLoader``"""

    def __init__(self, dataset, batch_size=1, shuffle=False, num_workers=0, pin_memory=True, timeout=0, **kwargs):
        super().__init__(dataset, batch_size=batch_size, shuffle=shuffle, **kwargs)


def split_indices(total_n, n_train, n_val, mode, generator=None):
    """reference ``dataloader.py:62-83``"""
    if isinstance(n_train, float):
        n_train = int(n_train * total_n)
    if isinstance(n_val, float):
        n_val = int(n_val * total_n)
    if n_train + n_val > total_n:
        raise ValueError("too little data for training and validation. please reduce n_train and n_val")
    if mode == "random":
        idcs = torch.randperm(total_n, generator=generator)
    elif mode == "sequential":
        idcs = torch.arange(total_n)
    else:
        raise NotImplementedError(f"splitting mode {mode} not implemented")
    return idcs[:n_train], idcs[n_train:n_train + n_val]


def rank_paths(paths, rank, world_size):
    """a list of dataset files is split among the processes (reference ``dataloader.py:56-60``)"""
    if not isinstance(paths, (list, tuple)):
        return paths
    gcd = math.gcd(world_size, len(paths))
    per = len(paths) // gcd
    return list(paths[(rank % gcd) * per:(rank % gcd + 1) * per])


def auto_reset(loader):
    while True:
        for batch in loader:
            yield batch


def getDataIters(config, rank=0, world_size=1, seed=0, device=None, dataset=None, **pipeline_kwargs):
    """-> (endless training iterator, endless validation iterator) of device batches (reference ``getDataIters``)"""
    dc = config.data_config
    if dataset is None:
        kwargs = dict(dc.items()) if hasattr(dc, "items") else dict(dc)
        kwargs["path"] = rank_paths(kwargs.get("path"), rank, world_size)
        dataset = CondensedDataset(**{k: v for k, v in kwargs.items()
                                      if k in ("path", "key_map", "type_names", "preprocess")})
    tr, va = split_indices(len(dataset), dc.n_train, dc.n_val, dc.train_val_split)
    rng = torch.Generator().manual_seed(seed + rank)      # every rank draws its own batches, as in the reference
    common = dict(batch_size=int(config.batch_size), drop_last=True, device=device, **pipeline_kwargs)
    train = DataLoader(dataset, indices=tr, shuffle=True, generator=rng, **common)
    val = DataLoader(dataset, indices=va, shuffle=False, **common)
    return train.endless(), val.endless()
