"""CUDA-graph execution of an evaluation step (B200 path: "CUDA streams and graphs instead of a tracing
compiler").  The interaction blocks issue ~120 small-to-medium launches per step; on a slow host the
Python / autograd bookkeeping around them costs more than the kernels.  ``GraphedEvaluator`` captures
model(batch) -- forward AND the position-gradient backward -- once per shape signature
(atoms, edges, graphs) and replays it; the neighbour list stays eager (it sizes the edge arrays, one
host synchronisation) and its result is copied into the graph's static buffers.

Shapes that never repeat (e.g. MD, where the number of edges changes every step) simply miss the cache and
run eagerly, so the evaluator is always safe to use; batched inference over fixed molecules and the
diffusion sampler (fixed complete graphs) hit it every time."""
import torch

from . import _lib, ops


class _Entry:
    __slots__ = ("graph", "static_in", "csr", "edge_index", "n_edges", "out", "launches")


class GraphedEvaluator:
    def __init__(self, model, r_max, attrs=None, out_keys=("energy", "forces"), max_entries=8):
        self.model, self.r_max, self.out_keys = model, float(r_max), tuple(out_keys)
        self.attrs = attrs or {"pos": ("node", "1x1o"), "species": ("node", "1x0e"), "_n_nodes": ("graph", "1x0e")}
        self.cache, self.max_entries = {}, max_entries
        self.hits = self.misses = 0

    def _weights_tag(self):
        """a captured graph has the packed weights of its capture baked in: new parameter values -> new capture"""
        return (ops.WEIGHTS_EPOCH, sum(p._version for p in self.model.parameters()))

    def _run(self, tensors, edge_index, n_edges):
        from e3_layers.data import Batch

        attrs = dict(self.attrs)
        attrs["_n_edges"] = ("graph", "1x0e")
        batch = Batch(attrs, edge_index=edge_index, _n_edges=n_edges, **tensors)
        out = self.model(batch)
        return {k: out[k] for k in self.out_keys}

    def __call__(self, tensors):
        """tensors: dict(pos [N,3] f32, species [N,1] i64, _n_nodes [G,1] i64, ...) on the GPU.
        -> dict of output tensors (owned by the evaluator until the next call with the same shapes)"""
        pos = tensors["pos"]
        st = ops.radius_graph_count(pos, tensors["_n_nodes"].reshape(-1), self.r_max)      # the step's one host sync
        key = (pos.shape[0], st.E, tensors["_n_nodes"].numel(), self._weights_tag()) + tuple(sorted(tensors))
        e = self.cache.get(key)
        if e is None:
            self.misses += 1
            edge_index, n_edges, csr = ops.radius_graph_finish(st)
            if len(self.cache) >= self.max_entries or self.model.training:
                return self._run(tensors, edge_index, n_edges)
            return self._capture(key, tensors, edge_index, n_edges, csr)
        self.hits += 1
        # hit: the GPU is idle from the synchronisation above until the replay below is enqueued, so as little as
        # possible is launched in between -- the edges are written straight into the graph's static buffers
        ops.radius_graph_fill(st, edge_index=e.edge_index, rev=e.csr.in_eid)
        e.csr.in_ptr.copy_(st.row_ptr, non_blocking=True)                 # out_ptr is the same tensor
        e.csr.in_nbr.copy_(e.edge_index[1], non_blocking=True)            # int64 -> int32
        e.n_edges.copy_(ops.edges_per_graph(st), non_blocking=True)
        for k, v in tensors.items():
            e.static_in[k].copy_(v, non_blocking=True)
        e.graph.replay()
        _lib.count_launch(e.launches)
        return e.out

    def _capture(self, key, tensors, edge_index, n_edges, csr):
        e = _Entry()
        e.static_in = {k: v.clone() for k, v in tensors.items()}
        e.edge_index, e.n_edges = edge_index.clone(), n_edges.clone()
        clone = lambda t: None if t is None else t.clone()
        in_ptr = clone(csr.in_ptr)
        # the radius graph shares row_ptr between the two groupings: keep that aliasing in the static copy
        out_ptr = in_ptr if csr.out_ptr is csr.in_ptr else clone(csr.out_ptr)
        e.csr = ops.GraphCSR(csr.n_nodes, csr.n_edges, in_ptr, clone(csr.in_nbr), clone(csr.in_eid), out_ptr, clone(csr.out_eid))
        e.edge_index._e3b_csr = e.csr
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up on the capture stream (lazy initialisation, packs)
            self._run(e.static_in, e.edge_index, e.n_edges)
        torch.cuda.current_stream().wait_stream(side)
        e.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count
        with torch.cuda.graph(e.graph):
            e.out = self._run(e.static_in, e.edge_index, e.n_edges)
        e.launches = _lib.launch_count - n0
        self.cache[key] = e
        e.graph.replay()
        _lib.count_launch(e.launches)
        return e.out
