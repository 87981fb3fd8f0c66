"""CUDA-graph execution of an evaluation step (B200 path: "CUDA streams and graphs instead of a tracing
compiler").  The interaction blocks issue ~120 small-to-medium launches per step; on a slow host the
Python / autograd bookkeeping around them costs more than the kernels.  ``GraphedEvaluator`` captures
model(batch) -- forward AND the position-gradient backward -- once per shape signature
(atoms, edges, graphs) and replays it; the neighbour list stays eager (it sizes the edge arrays, one
host synchronisation) and writes straight into the graph's static buffers.

**Bucketed shapes** (``node_bucket`` / ``edge_bucket`` > 0).  A stream of DIFFERENT batches almost never repeats an
exact (atoms, edges) pair, so the evaluator pads every batch up to the next bucket: padding atoms are appended as one
extra graph of isolated pairs, and the edge list is topped up with parallel copies of the pair edges
(``ops.radius_graph_fill_padded``), so that every array the captured kernels see has the bucket's shape and holds valid
finite data.  Real atoms keep exactly their own neighbours (graphs never share edges), so the energies and forces of the
real graphs are bit-identical to the unpadded evaluation; the outputs are sliced back to the real atoms / graphs.  The
padding work (a few per cent of the atoms and edges) is real work inside the timed step.

Without buckets the exact shapes are the key: batched inference over fixed molecules and the diffusion sampler (fixed
complete graphs) hit every time, anything else misses and runs eagerly.  Batches that bring their own edge list (complete
graphs, bond lists) or whose model builds it as a layer take the same route without any host synchronisation
(``_call_given_topology``; the model must then have static shapes for fixed inputs, which rules out a neighbour-list layer
with random pair criteria).  The evaluator is always safe to use: training
mode, edge-typed inputs or a full cache simply run the model eagerly."""
import torch

from . import _lib, ops


class _Entry:
    __slots__ = ("graph", "static_in", "csr", "edge_index", "n_edges", "out", "launches", "tag", "tick")


class _Stage:
    """padded input buffers of one (padded atoms, graphs) group; every capture of the group reads these"""
    __slots__ = ("tensors", "pad_pos")


def _round_up(x, m):
    return (x + m - 1) // m * m


class GraphedEvaluator:
    def __init__(self, model, r_max, attrs=None, out_keys=("energy", "forces"), max_entries=8, node_bucket=0,
                 edge_bucket=0, min_pad_nodes=128, grad=True):
        self.model, self.r_max, self.out_keys = model, float(r_max), tuple(out_keys)
        self.attrs = attrs or {"pos": ("node", "1x1o"), "species": ("node", "1x0e"), "_n_nodes": ("graph", "1x0e")}
        self.cache, self.max_entries = {}, max_entries
        self.node_bucket, self.edge_bucket, self.min_pad_nodes = int(node_bucket), int(edge_bucket), int(min_pad_nodes)
        if self.edge_bucket % 2:
            raise ValueError("edge_bucket must be even (the radius graph is symmetric)")
        self.bucketed = self.node_bucket > 0 and self.edge_bucket > 0
        self.grad = bool(grad)     # False: forward-only models (no autograd inside the step) run / are captured under no_grad
        self.stages = {}
        self.hits = self.misses = self.eager = 0
        self._params = list(model.parameters())
        self._pool = None          # one memory pool for all captures: they are replayed one at a time
        self._tick = 0

    def _weights_tag(self):
        """a captured graph has the packed weights of its capture baked in: new parameter values -> new capture"""
        return (ops.WEIGHTS_EPOCH, sum(p._version for p in self._params))

    def _run(self, tensors, edge_index, n_edges):
        from e3_layers.data import Batch

        attrs = dict(self.attrs)
        attrs["_n_edges"] = ("graph", "1x0e")
        batch = Batch(attrs, edge_index=edge_index, _n_edges=n_edges, **tensors)
        if self.grad:
            out = self.model(batch)
        else:
            with torch.no_grad():
                out = self.model(batch)
        return {k: out[k] for k in self.out_keys}

    # -- bucketed padding -----------------------------------------------------------------------
    def _kind(self, k):
        a = self.attrs.get(k)
        return a[0] if a is not None else None

    def _pad_inputs(self, tensors):
        """-> (stage, N, G, n_pad): the batch copied into the group's padded buffers, padding rows rewritten"""
        N, G = tensors["pos"].shape[0], tensors["_n_nodes"].numel()
        Nb = _round_up(N + self.min_pad_nodes, self.node_bucket)
        n_pad = Nb - N
        st = self.stages.get((Nb, G))
        dev = tensors["pos"].device
        if st is None:
            st = _Stage()
            st.tensors = {}
            for k, v in tensors.items():
                rows = Nb if self._kind(k) == "node" else G + 1
                st.tensors[k] = torch.zeros((rows,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
            st.pad_pos = ops.padding_positions(self.node_bucket + self.min_pad_nodes, self.r_max, dev)
            self.stages[(Nb, G)] = st
        for k, v in tensors.items():
            buf = st.tensors[k]
            if self._kind(k) == "node":
                buf[:N].copy_(v, non_blocking=True)
                if k == "pos":
                    buf[N:].copy_(st.pad_pos[:n_pad], non_blocking=True)
                else:
                    buf[N:].copy_(v[:1].expand(n_pad, *v.shape[1:]), non_blocking=True)   # any valid row
            else:
                buf[:G].copy_(v, non_blocking=True)
                if k == "_n_nodes":
                    buf[G:].fill_(n_pad)
                else:
                    buf[G:].copy_(v[:1], non_blocking=True)
        return st, N, G, n_pad

    def _slice(self, out, N, G, Nb):
        res = {}
        for k, v in out.items():
            if v.dim() and v.shape[0] == Nb:
                res[k] = v[:N]
            elif v.dim() and v.shape[0] == G + 1:
                res[k] = v[:G]
            else:
                res[k] = v
        return res

    # -- cache ----------------------------------------------------------------------------------
    def _lookup(self, key, tag):
        e = self.cache.get(key)
        if e is not None and e.tag != tag:           # stale weights: the capture is dead, free its slot
            del self.cache[key]
            e = None
        if e is not None:
            self._tick += 1
            e.tick = self._tick
        return e

    def _room(self, tag):
        """room for one more capture?  Entries captured with other weights are dead and go first.  A full cache of live
        entries: with buckets the least recently used one is replaced (the signatures are few and recur); with exact
        shapes nothing is evicted and the caller runs eagerly (a stream of never-repeating shapes, e.g. MD, must not pay a
        capture per step)."""
        for k in [k for k, e in self.cache.items() if e.tag != tag]:
            del self.cache[k]
        if len(self.cache) < self.max_entries:
            return True
        if not self.bucketed:
            return False
        victim = min(self.cache, key=lambda k: self.cache[k].tick)
        del self.cache[victim]
        return True

    def __call__(self, tensors):
        """tensors: dict(pos [N,3] f32, species [N,1] i64, _n_nodes [G,1] i64, ...) on the GPU.
        -> dict of output tensors (owned by the evaluator until the next call)"""
        eager = self.model.training or (self.grad and not torch.is_grad_enabled())
        if "edge_index" in tensors or self.r_max <= 0:
            return self._call_given_topology(tensors, eager)
        if self.bucketed and not eager and all(self._kind(k) in ("node", "graph") for k in tensors):
            return self._call_bucketed(tensors)
        pos = tensors["pos"]
        st = ops.radius_graph_count(pos, tensors["_n_nodes"].reshape(-1), self.r_max)      # the step's one host sync
        if eager:
            self.eager += 1
            edge_index, n_edges, _ = ops.radius_graph_finish(st)
            return self._run(tensors, edge_index, n_edges)
        tag = self._weights_tag()
        key = (pos.shape[0], st.E, tensors["_n_nodes"].numel()) + tuple(sorted(tensors))
        e = self._lookup(key, tag)
        if e is None:
            self.misses += 1
            edge_index, n_edges, csr = ops.radius_graph_finish(st)
            if not self._room(tag):
                self.eager += 1
                return self._run(tensors, edge_index, n_edges)
            static_in = {k: v.clone() for k, v in tensors.items()}
            return self._capture(key, tag, static_in, edge_index.clone(), n_edges.clone(), csr, clone_csr=True)
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

    def _call_given_topology(self, tensors, eager):
        """the batch brings its own edge list (complete graphs of the diffusion configs, bond lists) or the model builds
        its own (neighbour list as a model layer): no host synchronisation at all, one capture per input shape signature;
        the inputs are copied into the capture's static buffers"""
        if eager:
            self.eager += 1
            return self._run_plain(tensors)
        tag = self._weights_tag()
        key = tuple((k, tuple(v.shape)) for k, v in sorted(tensors.items()))
        e = self._lookup(key, tag)
        if e is None:
            self.misses += 1
            if not self._room(tag):
                self.eager += 1
                return self._run_plain(tensors)
            e = _Entry()
            e.tag = tag
            self._tick += 1
            e.tick = self._tick
            e.static_in = {k: v.clone() for k, v in tensors.items()}
            e.csr = e.edge_index = e.n_edges = None
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._run_plain(e.static_in)
            torch.cuda.current_stream().wait_stream(side)
            e.graph = torch.cuda.CUDAGraph()
            if self._pool is None or not self.cache:
                self._pool = torch.cuda.graph_pool_handle()
            n0 = _lib.launch_count
            with torch.cuda.graph(e.graph, pool=self._pool):
                e.out = self._run_plain(e.static_in)
            e.launches = _lib.launch_count - n0
            self.cache[key] = e
        else:
            self.hits += 1
            for k, v in tensors.items():
                e.static_in[k].copy_(v, non_blocking=True)
        e.graph.replay()
        _lib.count_launch(e.launches)
        return e.out

    def _run_plain(self, tensors):
        from e3_layers.data import Batch

        batch = Batch(dict(self.attrs), **dict(tensors))        # (a fresh dict: the model's layers add their keys to it)
        if self.grad:
            out = self.model(batch)
        else:
            with torch.no_grad():
                out = self.model(batch)
        return {k: out[k] for k in self.out_keys}

    def _call_bucketed(self, tensors):
        stage, N, G, n_pad = self._pad_inputs(tensors)
        Nb = N + n_pad
        padded = stage.tensors
        st = ops.radius_graph_count(padded["pos"], padded["_n_nodes"].reshape(-1), self.r_max)   # the one host sync
        Eb = _round_up(st.E, self.edge_bucket)
        tag = self._weights_tag()
        key = (Nb, Eb, G) + tuple(sorted(tensors))
        e = self._lookup(key, tag)
        if e is not None:
            self.hits += 1
            ops.radius_graph_fill_padded(st, n_pad, Eb, e.edge_index, e.csr.in_eid, e.csr.in_nbr)
            e.csr.in_ptr.copy_(st.row_ptr, non_blocking=True)
            e.n_edges.copy_(ops.edges_per_graph(st), non_blocking=True)
            e.graph.replay()
            _lib.count_launch(e.launches)
            return self._slice(e.out, N, G, Nb)
        self.misses += 1
        dev = padded["pos"].device
        edge_index = torch.empty(2, Eb, dtype=torch.int64, device=dev)
        rev = torch.empty(Eb, dtype=torch.int32, device=dev)
        nbr = torch.empty(Eb, dtype=torch.int32, device=dev)
        ops.radius_graph_fill_padded(st, n_pad, Eb, edge_index, rev, nbr)
        row_ptr = st.row_ptr.clone()
        csr = ops.GraphCSR(Nb, Eb, row_ptr, nbr, rev, row_ptr, None)
        n_edges = ops.edges_per_graph(st).clone()
        self._room(tag)
        out = self._capture(key, tag, padded, edge_index, n_edges, csr, clone_csr=False)
        return self._slice(out, N, G, Nb)

    def _capture(self, key, tag, static_in, edge_index, n_edges, csr, clone_csr):
        e = _Entry()
        e.tag = tag
        self._tick += 1
        e.tick = self._tick
        e.edge_index, e.n_edges = edge_index, n_edges
        if clone_csr:
            clone = lambda t: None if t is None else t.clone()
            in_ptr = clone(csr.in_ptr)
            # the radius graph shares row_ptr between the two groupings: keep that aliasing in the static copy
            out_ptr = in_ptr if csr.out_ptr is csr.in_ptr else clone(csr.out_ptr)
            csr = ops.GraphCSR(csr.n_nodes, csr.n_edges, in_ptr, clone(csr.in_nbr), clone(csr.in_eid), out_ptr, clone(csr.out_eid))
        e.csr = csr
        e.edge_index._e3b_csr = e.csr
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up on the capture stream (lazy initialisation, packs)
            self._run(static_in, e.edge_index, e.n_edges)
        torch.cuda.current_stream().wait_stream(side)
        e.graph = torch.cuda.CUDAGraph()
        if self._pool is None or not self.cache:       # the allocator retires a pool with its last graph
            self._pool = torch.cuda.graph_pool_handle()
        n0 = _lib.launch_count
        with torch.cuda.graph(e.graph, pool=self._pool):
            e.out = self._run(static_in, e.edge_index, e.n_edges)
        e.launches = _lib.launch_count - n0
        e.static_in = static_in
        self.cache[key] = e
        e.graph.replay()
        _lib.count_launch(e.launches)
        return e.out
