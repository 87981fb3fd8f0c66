"""Neighbour list and edge vectors on the B200 (reference ``e3_layers/data/compute_edge.py``).

``computeEdgeIndex`` keeps the reference contract -- returns ``({"edge_index": [2,E] int64}, attrs)``
with edges in lexicographic (source, destination) order, writes ``_n_edges`` into the incoming
``data`` dict -- but runs the all-pairs-within-a-graph radius search as a CUDA kernel
(``e3b_radius_graph_*``) with the reference's exact fp32 predicate.  ``criteria`` edges are OR-ed
in from the callable's mask, evaluated only on the candidate pairs it can add."""
import torch

from e3b200 import ops


def computeEdgeVector(data, attrs, key="pos", with_lengths=True):
    attrs["edge_vector"] = ("edge", "1x1o")
    attrs["edge_length"] = ("edge", "1x0e")
    if "edge_vector" in data:
        if with_lengths and "edge_length" not in data:
            data["edge_length"] = torch.linalg.norm(data["edge_vector"], dim=-1)
        return data, attrs
    pos, ei = data[key], data["edge_index"]
    csr = ops.graph_of(ei, pos.shape[0])
    vec, length = ops.edge_vectors(pos, ei, csr)
    data["edge_vector"] = vec
    if with_lengths:
        data["edge_length"] = length
    return data, attrs


def _all_pairs(n_nodes, device):
    """[2, sum_g n_g^2]: per graph every ordered pair (a, b), a slow, b fast (reference order)"""
    counts = n_nodes.reshape(-1).to(device)
    size = torch.repeat_interleave(counts, counts)                      # [N] size of each node's graph
    first = torch.repeat_interleave(torch.cumsum(counts, 0) - counts, counts)   # [N] first node of that graph
    a = torch.repeat_interleave(torch.arange(size.numel(), device=device), size)
    row_start = torch.cumsum(size, 0) - size                            # [N] first pair index of node a
    b = first[a] + (torch.arange(a.numel(), device=device) - row_start[a])
    return torch.stack([a, b])


def computeEdgeIndex(data, attrs, r_max=None, key="pos", criteria=None):
    pos = data[key]
    if not pos.is_cuda:
        raise RuntimeError("computeEdgeIndex (B200 path) needs CUDA tensors; there is no CPU fallback")
    n_nodes = data["_n_nodes"].reshape(-1)
    if "edge_index" in data:
        raise NotImplementedError("merging a pre-existing edge_index (compute_edge.py:77-100) is out of scope; "
                                  "drop 'edge_index' before recomputing, as sde_sampling.py:237-242 does")
    edge_index, n_edges, csr = ops.radius_graph(pos, n_nodes, r_max)
    if criteria is not None:
        pairs = _all_pairs(n_nodes, pos.device)
        extra = criteria(data, pairs) & (pairs[0] != pairs[1])
        N = pos.shape[0]
        keys = torch.cat([edge_index[0] * N + edge_index[1], (pairs[0] * N + pairs[1])[extra]])
        keys = torch.unique(keys)                                # sorted -> reference order
        edge_index = torch.stack([keys // N, keys % N])
        seg = torch.repeat_interleave(torch.arange(n_nodes.numel(), device=pos.device), n_nodes.to(pos.device))
        n_edges = torch.bincount(seg[edge_index[0]], minlength=n_nodes.numel()).view(-1, 1)
    attrs["_n_edges"] = ("graph", "1x0e")
    data["_n_edges"] = n_edges
    return {"edge_index": edge_index}, attrs
