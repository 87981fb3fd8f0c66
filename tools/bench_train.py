#!/usr/bin/env python3
"""Times the force-matching TRAINING step of config_energy_force at W2 (512 synthetic QM9-shaped molecules
per GPU): neighbour list, forward, position gradient with its graph (second-order mode), the reference's
loss 1e3 MSE(E) + 3e4 MSE(F), backward to the parameters, flat gradient all-reduce (N > 1) and Adam.

  python tools/bench_train.py [--graphs 512] [--steps 5] [--warmup 2]
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_train.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "equivariant-nn-zoo_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graphs", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--profile", default=None, help="write a torch profiler table of 3 steps to this file")
    a = ap.parse_args()
    import product_harness
    from e3_layers.data import Batch, computeEdgeIndex
    from e3b200 import _lib, optim, parallel, synthetic

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = product_harness.build_product({"config": "config_energy_force", "seed": 0}, torch.float32, dev).train()
    parallel.broadcast_parameters(model)
    opt = optim.FlatAdam(model, lr=1e-3)
    host = synthetic.qm9_like(a.graphs, seed=rank)
    attrs = {"pos": ("node", "1x1o"), "species": ("node", "1x0e"), "_n_nodes": ("graph", "1x0e")}
    res = {k: v.to(dev) for k, v in host.items()}
    n_atoms = host["pos"].shape[0]
    g = torch.Generator().manual_seed(1)
    e_t = torch.randn(a.graphs, 1, generator=g).to(dev)
    f_t = (0.1 * torch.randn(n_atoms, 3, generator=g)).to(dev)

    def step():
        batch = Batch(dict(attrs), **{k: v.clone() for k, v in res.items()})
        d, at = computeEdgeIndex(batch.data, batch.attrs, r_max=5.0)
        batch.update(d)
        batch.attrs.update(at)
        out = model(Batch(batch.attrs, **batch.data))
        loss = 1e3 * ((out["energy"] - e_t) ** 2).mean() + 3e4 * ((out["forces"] - f_t) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.all_reduce()
        opt.step()
        return loss

    losses = []
    for _ in range(a.warmup):
        losses.append(float(step().detach()))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.reset_peak_memory_stats()
    n0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    losses.append(float(loss.detach()))
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    atoms = torch.tensor([float(n_atoms)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(atoms)
    if rank == 0:
        ms = float(t) / a.steps
        print(json.dumps({"what": "config_energy_force force-matching training step (second-order mode), W2",
                          "n_gpus": world, "graphs_per_gpu": a.graphs, "atoms_per_gpu": n_atoms, "ms_per_step": ms,
                          "train_atoms_per_s": float(atoms) / (ms * 1e-3), "libe3b200_launches_per_step":
                          (_lib.launch_count - n0) / a.steps, "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
                          "losses": losses}))
    if a.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
        with open(a.profile, "w") as f:
            f.write("# torch profiler, 3 training steps of config_energy_force at W2 (tools/bench_train.py --profile)\n")
            f.write(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=70))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
