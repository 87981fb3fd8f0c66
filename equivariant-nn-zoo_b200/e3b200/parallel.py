"""Data parallelism of the hot path (reference: ``train.py:272,280-304`` spawns one process per GPU and
wraps the model in DistributedDataParallel, ``run/trainer.py:138-139``).

Graphs are independent units (``Batch`` is a concatenation without cross-graph edges), so the work is
sharded by graph, balanced by edge count; inference needs no collective.  Training adds ONE all-reduce
of a flat fp32 gradient buffer per optimiser step (NCCL over NVLink on the GPU box, gloo in the CPU
tests) instead of DDP's per-bucket hooks; scalars for logging ride in the tail of the same buffer."""
import torch
import torch.distributed as dist


def shard_graphs(cost, world_size):
    """Greedy longest-processing-time bin packing.  cost: per-graph work estimate (e.g. n (n - 1), the
    number of candidate edges).  -> list (one per rank) of graph index lists, each in ascending order;
    deterministic, identical on every rank."""
    cost = [float(c) for c in cost]
    order = sorted(range(len(cost)), key=lambda i: (-cost[i], i))
    loads = [0.0] * world_size
    bins = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        bins[r].append(i)
        loads[r] += cost[i]
    return [sorted(b) for b in bins]


def shard_batch(tensors, rank, world_size):
    """The graphs of rank `rank` from a host-side dict with `_n_nodes` [G,1] and per-node tensors
    (everything with leading dimension N); per-graph tensors (leading dimension G) are indexed too."""
    n = tensors["_n_nodes"].reshape(-1)
    G, N = n.numel(), int(n.sum())
    mine = shard_graphs((n * (n - 1)).tolist(), world_size)[rank]
    starts = torch.cumsum(n, 0) - n
    node_idx = torch.cat([torch.arange(int(starts[g]), int(starts[g] + n[g])) for g in mine]) if mine else torch.zeros(0, dtype=torch.long)
    g_idx = torch.tensor(mine, dtype=torch.long)
    out = {}
    for k, v in tensors.items():
        if v.shape[0] == N and k != "_n_nodes":
            out[k] = v[node_idx]
        elif v.shape[0] == G:
            out[k] = v[g_idx]
        else:
            raise ValueError(f"cannot shard {k} with shape {tuple(v.shape)}")
    return out


class FlatGradients:
    """Flat fp32 gradient buffer over the trainable parameters: one all-reduce (sum, then / world) per step."""

    def __init__(self, params, n_scalars=0):
        self.params = [p for p in params if p.requires_grad]
        self.sizes = [p.numel() for p in self.params]
        total = sum(self.sizes)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.buf = torch.zeros(total + n_scalars, dtype=torch.float32, device=dev)
        self.n_scalars = n_scalars
        self.total = total

    def all_reduce(self, scalars=None):
        """averages the gradients over the ranks in place; returns the averaged scalars"""
        o = 0
        for p, n in zip(self.params, self.sizes):
            if p.grad is None:
                self.buf[o:o + n].zero_()
            else:
                self.buf[o:o + n].copy_(p.grad.reshape(-1))
            o += n
        if self.n_scalars:
            self.buf[o:] = torch.as_tensor(scalars if scalars is not None else [0.0] * self.n_scalars,
                                           dtype=torch.float32, device=self.buf.device)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM)
            self.buf /= dist.get_world_size()
        o = 0
        for p, n in zip(self.params, self.sizes):
            if p.grad is None:
                p.grad = self.buf[o:o + n].view_as(p).clone()
            else:
                p.grad.copy_(self.buf[o:o + n].view_as(p))
            o += n
        return self.buf[o:].clone() if self.n_scalars else None


def broadcast_parameters(module, src=0):
    """identical initial weights on every rank (the reference seeds every rank alike, train.py:59,97)"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src)
