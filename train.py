#!/usr/bin/env python3
"""Training entry point of the B200 drop-in (reference ``train.py``: same flag names for the in-scope
options, ``--config <name>`` picks a registered config, one process per GPU, rank 0 logs and saves).

In scope: regression on energies / dipoles (first-order parameter gradients through the fused
interaction blocks) and energy+force matching (``config_energy_force``: the graph of the position
gradient is built by the second-order mode of ``GradientOutput``), data-parallel over graphs with ONE
flat gradient all-reduce per step, and score matching for the diffusion configs (``train_diffusion``: VP-SDE
loss of ``e3_layers.run``, EMA, gradient clipping / accumulation).
Data: ``--data synthetic`` (seeded QM9-shaped molecules with a synthetic per-species target; there is no
network for datasets) or an ``.npz`` with ``pos, species, _n_nodes`` and the target key.

  python train.py --config config_energy --steps 20
  python train.py --config config_energy --world_size 8        # spawns one process per GPU (reference launch_mp)
"""
import argparse
import logging
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "equivariant-nn-zoo_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--config_spec", default="")
    ap.add_argument("--name", default="default")
    ap.add_argument("--workdir", default="results")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--resume_from", default=None)
    ap.add_argument("--log_period", type=int, default=10)
    ap.add_argument("--save_period", type=int, default=2000)
    ap.add_argument("--world_size", type=int, default=1)
    ap.add_argument("--master_addr", default="127.0.0.1")
    ap.add_argument("--master_port", default="10000")
    ap.add_argument("--verbose", default="INFO")
    ap.add_argument("--steps", type=int, default=100, help="optimiser steps (the reference runs until early stopping)")
    ap.add_argument("--data", default="synthetic", help="'synthetic' or the path of an .npz dataset")
    ap.add_argument("--n_graphs", type=int, default=4096, help="size of the synthetic dataset")
    # score-based configs: the reference reads these from a score_sde_pytorch config file (--sde_config, an external
    # repository); the defaults below are that repository's configs/vp/cifar10_ncsnpp_continuous.py, which the
    # reference's README uses
    ap.add_argument("--sde_config", default=None, help="accepted for compatibility; the VP-SDE settings come from the flags below")
    ap.add_argument("--beta_min", type=float, default=0.1)
    ap.add_argument("--beta_max", type=float, default=20.0)
    ap.add_argument("--num_scales", type=int, default=1000)
    ap.add_argument("--ema_rate", type=float, default=0.9999)
    ap.add_argument("--n_res", type=int, default=400, help="residues of the synthetic protein graphs (config_diffusion_CA)")
    return ap.parse_args()


def load_data(flags, config, target_keys):
    from e3b200 import synthetic

    if flags.data != "synthetic":
        z = np.load(flags.data)
        return {k: torch.from_numpy(z[k]) for k in z.files}
    data = synthetic.qm9_like(flags.n_graphs, seed=flags.seed)
    n = data["_n_nodes"].reshape(-1)
    seg = torch.repeat_interleave(torch.arange(n.numel()), n)
    gen = torch.Generator().manual_seed(flags.seed + 1)
    if "dipole" in target_keys:                                  # per-node 1x1o target
        data["dipole"] = 0.1 * torch.randn(data["pos"].shape[0], 3, generator=gen)
    energy_key = next((k for k in target_keys if k in ("energy", "total_energy")), None)
    if energy_key is not None:
        # per-graph scalar: composition energy + a smooth pair potential, so that energies and forces are consistent
        per_species = -torch.arange(0, 120, dtype=torch.float32) * 0.37
        pos = data["pos"].double().requires_grad_(True)
        starts = torch.cumsum(n, 0) - n
        ii, jj = [], []
        for g in range(n.numel()):
            a = torch.arange(int(starts[g]), int(starts[g] + n[g]))
            i, j = torch.meshgrid(a, a, indexing="ij")
            keep = i < j
            ii.append(i[keep])
            jj.append(j[keep])
        ii, jj = torch.cat(ii), torch.cat(jj)
        r = (pos[ii] - pos[jj]).norm(dim=-1)
        pair = torch.zeros(n.numel(), dtype=torch.float64).index_add_(0, seg[ii], 0.5 * torch.exp(-r))
        e = torch.zeros(n.numel(), dtype=torch.float64).index_add_(0, seg, per_species[data["species"].reshape(-1)].double()) + pair
        if "forces" in target_keys:
            (g_pos,) = torch.autograd.grad(e.sum(), pos)
            data["forces"] = (-g_pos).float()
        data[energy_key] = (e.detach().float() + 0.01 * torch.randn(n.numel(), generator=gen)).view(-1, 1)
    return data


ATTRS = {"pos": ("node", "1x1o"), "species": ("node", "1x0e"), "_n_nodes": ("graph", "1x0e"),
         "energy": ("graph", "1x0e"), "total_energy": ("graph", "1x0e"), "forces": ("node", "1x1o"), "dipole": ("node", "1x1o")}


def make_pipeline(data, target_keys, batch_size, r_max, seed, rank, world, dev, resident=None):
    """dataset -> device (e3_layers.data.DevicePipeline): the concatenated host tensors go to HBM once (or stay pinned
    and are streamed by the copy engine), every batch is cut out by vectorised gathers and the neighbour list is built
    on the GPU as a preprocess with the layer contract -- no DataLoader workers (reference dataloader.py:86-105)"""
    from functools import partial

    from e3_layers.data import CondensedDataset, DevicePipeline, computeEdgeIndex

    attrs = {k: ATTRS[k] for k in data if k in ATTRS}
    ds = CondensedDataset(data=data, attrs=attrs, preprocess=[partial(computeEdgeIndex, r_max=r_max)])
    gen = torch.Generator().manual_seed(seed)                     # same order on every rank; each takes its shard
    batch_size = max(1, min(batch_size, int(data["_n_nodes"].shape[0]) // world))    # a dataset smaller than one batch
    return DevicePipeline(ds, batch_size=batch_size, shuffle=True, drop_last=True, generator=gen, device=dev,
                          rank=rank, world_size=world, resident=resident)


def train_step(model, opt, batch, target_keys, loss_coeffs):
    """one optimiser step on a device batch that carries its targets; -> (loss, mae) as device scalars"""
    from e3_layers.data import Batch

    targets = {k: batch.data.pop(k) for k in target_keys}
    for k in target_keys:
        batch.attrs.pop(k, None)
    out = model(Batch(batch.attrs, **batch.data))
    loss, mae = 0.0, 0.0                                         # reference run/loss.py:274-287: sum of coeff * mean loss
    for k in target_keys:
        coeff, kind = loss_coeffs[k][0], loss_coeffs[k][1]
        diff = out[k] - targets[k]
        loss = loss + coeff * (diff.abs().mean() if kind == "L1Loss" else (diff ** 2).mean())
        mae = mae + diff.detach().abs().mean()
    opt.zero_grad()
    loss.backward()
    opt.all_reduce()
    opt.step()
    return loss.detach(), mae


def main(rank, flags):
    from e3_layers import configs
    from e3_layers.utils import build, setSeed
    from e3b200 import optim, parallel

    world = flags.world_size
    if "RANK" in os.environ:                                     # launched by torchrun
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        local = int(os.environ.get("LOCAL_RANK", rank))
    else:
        local = rank
        os.environ.setdefault("MASTER_ADDR", flags.master_addr)
        os.environ.setdefault("MASTER_PORT", flags.master_port)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    logging.basicConfig(level=getattr(logging, flags.verbose), format=f"[rank {rank}] %(message)s")

    get = getattr(configs, flags.config, None)
    assert get is not None, f"Config {flags.config} not found."
    config = get(flags.config_spec) if flags.config_spec else get()
    if "diffusion" in flags.config:
        return train_diffusion(config, flags, rank, world, dev)
    loss_coeffs = dict(config.loss_coeffs.items()) if hasattr(config.loss_coeffs, "items") else dict(config.loss_coeffs)
    target_keys = list(loss_coeffs)
    setSeed(flags.seed)                                          # identical initial weights on every rank
    model = build(config.model_config).to(dev).train()
    resume = _load_checkpoint(flags.resume_from, dev)
    if resume is not None:
        model.load_state_dict(_strip_module(resume["model"]))
    parallel.broadcast_parameters(model)
    # flat parameter / gradient buffers: one all-reduce, one fused Adam + EMA kernel per step (e3b200.optim)
    opt = optim.FlatAdam(model, lr=float(config.learning_rate),
                         ema_decay=float(config.ema_decay) if getattr(config, "use_ema", False) else None,
                         ema_use_num_updates=bool(getattr(config, "ema_use_num_updates", True)))
    first_step = 0
    if resume is not None and resume.get("optimizer") is not None:     # full trainer state: moments, EMA, progress
        opt.load_state_dict(resume["optimizer"])
        first_step = int(resume.get("step", 0))
    r_max = float(config.model_config.r_max)

    data = load_data(flags, config, target_keys)
    pipe = make_pipeline(data, target_keys, int(config.batch_size), r_max, flags.seed, rank, world, dev)
    batches = pipe.endless(skip=first_step)                      # a resumed run replays the data order up to its step
    out_dir = os.path.join(flags.workdir, flags.name)
    t0 = time.time()
    for step in range(first_step, flags.steps):
        loss, mae = train_step(model, opt, next(batches), target_keys, loss_coeffs)
        if step % flags.log_period == 0 or step == flags.steps - 1:
            scal = torch.stack([loss, torch.as_tensor(mae, device=dev)])
            if world > 1:                                        # logging scalars: averaged over the ranks on log steps only
                dist.all_reduce(scal)
                scal /= world
            if rank == 0:
                logging.info("step %d loss %.6g mae %.6g (%.1f s)", step, float(scal[0]), float(scal[1]), time.time() - t0)
        if rank == 0 and ((step + 1) % flags.save_period == 0 or step == flags.steps - 1):
            # reference Trainer.save (run/trainer.py:632-763): trainer.pt = everything needed to resume (raw weights,
            # optimiser moments, EMA, progress); model.pt = the deployable weights = the EMA average when it is kept
            os.makedirs(out_dir, exist_ok=True)
            _atomic_save({"model": model.state_dict(), "optimizer": opt.state_dict(), "step": step + 1,
                          "config": flags.config}, os.path.join(out_dir, "trainer.pt"))
            _atomic_save(opt.ema_state_dict(model), os.path.join(out_dir, "model.pt"))
    if world > 1:
        dist.destroy_process_group()


def _strip_module(sd):
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def _load_checkpoint(path, dev):
    """--resume_from takes a trainer.pt (full state) or a bare state_dict (weights only, reference inference.py:46-53)"""
    if not path:
        return None
    state = torch.load(path, map_location=dev, weights_only=False)
    if isinstance(state, dict) and "model" in state and isinstance(state["model"], dict):
        return state
    return {"model": state}


def _atomic_save(obj, path):
    tmp = path + ".tmp"
    torch.save(obj, tmp)
    os.replace(tmp, path)                                        # the reference renames a finished temp file too


def train_diffusion(config, flags, rank, world, dev):
    """Score-matching training (reference train.py:70-215 ``train_diffusion``): VP-SDE perturbation of the diffused
    keys, continuous-time loss, gradient clipping / accumulation from the e3 config, EMA of the parameters, one flat
    gradient all-reduce per step.  Synthetic data: W4 molecules (config_diffusion) or W5 C-alpha graphs."""
    from e3_layers.data import Batch
    from e3_layers.run import VPSDE, ExponentialMovingAverage, get_step_fn
    from e3_layers.utils import build, setSeed
    from e3b200 import parallel, synthetic

    setSeed(flags.seed)
    model = build(config.model_config).to(dev).train()
    resume = _load_checkpoint(flags.resume_from, dev)
    if resume is not None:
        model.load_state_dict(_strip_module(resume["model"]))
    parallel.broadcast_parameters(model)
    keys = getattr(config, "diffusion_keys", None)
    keys = dict(keys.items()) if keys is not None and hasattr(keys, "items") else {"pos": 3}
    sde = VPSDE(keys, beta_min=flags.beta_min, beta_max=flags.beta_max, N=flags.num_scales)
    opt = torch.optim.Adam(model.parameters(), lr=float(config.learning_rate))
    flat = parallel.FlatGradients(model.parameters())
    state = {"model": model, "optimizer": opt, "ema": ExponentialMovingAverage(model.parameters(), decay=flags.ema_rate), "step": 0}
    if resume is not None and resume.get("optimizer") is not None:     # reference restore_checkpoint: optimizer, ema, step
        opt.load_state_dict(resume["optimizer"])
        state["ema"].load_state_dict(resume["ema"])
        state["step"] = int(resume["step"])
    clip = getattr(config, "grad_clid_norm", None)
    step_fn = get_step_fn(sde, train=True, optimizer=opt, reduce_mean=True, grad_clid_norm=clip,
                          grad_acc=int(getattr(config, "grad_acc", 1)), grad_sync=flat.all_reduce if world > 1 else None)
    protein = "CA" in keys
    attrs_of = {"pos": ("node", "1x1o"), "CA": ("node", "1x1o"), "species": ("node", "1x0e"), "chain_id": ("node", "1x0e"),
                "id": ("node", "1x0e"), "t": ("graph", "1x0e"), "bond_type": ("edge", "1x0e"), "_n_nodes": ("graph", "1x0e"),
                "_n_edges": ("graph", "1x0e")}
    pool = []
    for i in range(4):                                           # a small pool of synthetic batches, one shard per rank
        seed = flags.seed + 100 * i + rank
        b = synthetic.protein_like(flags.n_res, seed=seed) if protein else synthetic.diffusion_like(int(config.batch_size), seed=seed)
        if protein:                                              # the model's first layer builds the neighbour list itself
            b.pop("edge_index"), b.pop("_n_edges")
        b.pop("t")
        pool.append({k: v.to(dev) for k, v in b.items()})
    out_dir = os.path.join(flags.workdir, flags.name)
    t0, window = time.time(), []
    for step in range(state["step"], flags.steps):
        b = pool[step % len(pool)]
        batch = Batch({k: attrs_of[k] for k in b if k in attrs_of}, **{k: v.clone() for k, v in b.items()})
        loss, losses = step_fn(state, batch)
        window.append(loss)
        if rank == 0 and (step % flags.log_period == 0 or step == flags.steps - 1):
            logging.info("step %d training_loss %.5e (%.1f s)", step, sum(window) / len(window), time.time() - t0)
            window = []
        if rank == 0 and ((step + 1) % flags.save_period == 0 or step == flags.steps - 1):
            # reference save_checkpoint(state) (utils/saveload.py:432-454): {optimizer, model, ema, step}; model.pt holds
            # the EMA weights, which are the ones the reference samples and validates with
            os.makedirs(out_dir, exist_ok=True)
            _atomic_save({"model": model.state_dict(), "optimizer": opt.state_dict(), "ema": state["ema"].state_dict(),
                          "step": state["step"], "config": flags.config}, os.path.join(out_dir, "checkpoint.pt"))
            ema = state["ema"]
            ema.store(model.parameters())
            ema.copy_to(model.parameters())
            _atomic_save(model.state_dict(), os.path.join(out_dir, "model.pt"))
            ema.restore(model.parameters())
    if world > 1:
        dist.destroy_process_group()


def launch_mp(flags):
    if flags.world_size > 1 and "RANK" not in os.environ:
        mp.spawn(main, args=(flags,), nprocs=flags.world_size, join=True)
    else:
        main(0, flags)


if __name__ == "__main__":
    launch_mp(parse())
