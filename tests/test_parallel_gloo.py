"""Host logic of the multi-GPU path on CPU: graph sharding and the flat gradient all-reduce
(world_size 2, gloo; the GPU box runs the same code over NCCL)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from e3b200 import parallel, synthetic


def test_shard_graphs_balanced_and_complete():
    n = torch.randint(3, 30, (257,), generator=torch.Generator().manual_seed(0))
    cost = (n * (n - 1)).tolist()
    for world in (1, 2, 4, 8):
        bins = parallel.shard_graphs(cost, world)
        assert sorted(i for b in bins for i in b) == list(range(257))
        loads = [sum(cost[i] for i in b) for b in bins]
        assert max(loads) - min(loads) <= max(cost)          # LPT bound
    assert parallel.shard_graphs([], 2) == [[], []]


def test_shard_batch_keeps_graphs_whole():
    host = synthetic.qm9_like(19, seed=3)
    parts = [parallel.shard_batch(host, r, 2) for r in range(2)]
    assert sum(p["pos"].shape[0] for p in parts) == host["pos"].shape[0]
    assert sum(p["_n_nodes"].numel() for p in parts) == 19
    for p in parts:
        assert int(p["_n_nodes"].sum()) == p["pos"].shape[0] == p["species"].shape[0]
    # every molecule's coordinates arrive intact on exactly one rank
    key = lambda t: sorted(round(float(x), 5) for x in t.reshape(-1))
    all_rows = key(torch.cat([p["pos"] for p in parts]))
    assert all_rows == key(host["pos"])


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        lin = torch.nn.Linear(5, 3)
        frozen = torch.nn.Parameter(torch.ones(2), requires_grad=False)
        parallel.broadcast_parameters(lin)
        x = torch.full((4, 5), float(rank + 1))
        lin(x).sum().backward()
        local = [p.grad.clone() for p in lin.parameters()]
        flat = parallel.FlatGradients(list(lin.parameters()) + [frozen], n_scalars=2)
        scal = flat.all_reduce([float(rank), 10.0])
        q.put((rank, [g.tolist() for g in local], [p.grad.tolist() for p in lin.parameters()], scal.tolist()))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, g0, s0), (_, l1, g1, s1) = res
    assert g0 == g1                                            # both ranks hold the same averaged gradient
    for a, b, avg in zip(l0, l1, g0):
        exp = ((torch.tensor(a) + torch.tensor(b)) / 2).tolist()
        assert torch.allclose(torch.tensor(avg), torch.tensor(exp))
    assert s0 == s1 == [0.5, 10.0]


def _worker_flat_adam(rank, world, port, q):
    from e3b200 import optim

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        lin = torch.nn.Linear(5, 3)
        parallel.broadcast_parameters(lin)
        opt = optim.FlatAdam(lin, lr=1e-2)                       # parameters and gradients become views of flat buffers
        x = torch.full((4, 5), float(rank + 1))
        opt.zero_grad()
        lin(x).sum().backward()
        aliased = all(p.grad.data_ptr() >= opt.grad.data_ptr() for p in lin.parameters())
        local = opt.grad.clone()
        opt.all_reduce()                                         # ONE collective on the buffer, no gather / scatter copies
        loud = False
        try:
            opt.step()                                           # the fused kernel needs a CUDA device: no CPU fallback
        except RuntimeError:
            loud = True
        q.put((rank, local.tolist(), opt.grad.tolist(), [p.grad.reshape(-1).tolist() for p in lin.parameters()], aliased, loud))
    finally:
        dist.destroy_process_group()


def test_flat_adam_buffers_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29411 + os.getpid() % 200
    procs = [ctx.Process(target=_worker_flat_adam, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, g0, v0, a0, loud0), (_, l1, g1, v1, a1, loud1) = res
    assert a0 and a1 and loud0 and loud1
    assert g0 == g1 and torch.allclose(torch.tensor(g0), (torch.tensor(l0) + torch.tensor(l1)) / 2)
    assert sum(v0, []) == g0                                     # the parameters' .grad ARE the buffer


def _worker_overlap(rank, world, port, q):
    from e3b200 import optim

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 7), torch.nn.Tanh(), torch.nn.Linear(7, 5), torch.nn.Tanh(),
                                  torch.nn.Linear(5, 2))
        unused = torch.nn.Linear(3, 3)                           # parameters that never receive a gradient
        holder = torch.nn.ModuleList([net, unused])
        parallel.broadcast_parameters(holder)
        opt = optim.FlatAdam(holder, lr=1e-2)
        opt.enable_overlap(bucket_bytes=128)                     # several buckets, all-reduce issued from gradient hooks
        res = []
        for it in range(2):                                      # second pass: counters were reset
            x = torch.full((4, 6), float(rank + 1 + it))
            opt.zero_grad()
            net(x).sum().backward()
            issued = sum(b["work"] is not None for b in opt._buckets)
            opt.all_reduce()
            res.append((issued, opt.grad.tolist()))
        # reference: the same two passes, plain all-reduce of the local gradient
        ref = []
        for it in range(2):
            x = torch.full((4, 6), float(rank + 1 + it))
            opt.zero_grad()
            for b in opt._buckets:
                b["n"] = 10 ** 9                                 # hooks never complete a bucket
            net(x).sum().backward()
            assert all(b["work"] is None for b in opt._buckets)
            opt.all_reduce()
            ref.append(opt.grad.tolist())
        q.put((rank, len(opt._buckets), res, ref))
    finally:
        dist.destroy_process_group()


def test_flat_adam_overlapped_allreduce_world2():
    """bucketed all-reduce issued from post-accumulate hooks during backward == one all-reduce after it"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29811 + os.getpid() % 200
    procs = [ctx.Process(target=_worker_overlap, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, nb0, r0, ref0), (_, nb1, r1, ref1) = res
    assert nb0 == nb1 and nb0 >= 3
    for (issued0, g0), (issued1, g1), e0 in zip(r0, r1, ref0):
        assert issued0 == issued1 and 1 <= issued0 < nb0          # some buckets went during backward, the unused one after
        assert g0 == g1
        assert torch.allclose(torch.tensor(g0), torch.tensor(e0))
    assert ref0 == ref1
