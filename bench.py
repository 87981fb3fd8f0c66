#!/usr/bin/env python3
"""Benchmark of the B200 hot path (BASELINE.json metric: energy+force atoms/s; fused TP-conv
edges/s and % of the HBM roofline).

Default workload W2 (BASELINE configs[1]): a "step" = one energy+force evaluation of the reference's
``config_energy_force`` model (n_dim 64, l_max 2, 5 interaction blocks, r_max 5) on one batch of 512 synthetic
QM9-shaped molecules per GPU: neighbour list + forward + position-gradient backward, through the public ``e3_layers``
API (``GradientOutput.forward``).  The timed region ROTATES over 8 different seeded batches per rank (different atom
and edge counts); the model step is a CUDA graph replayed over bucket-padded shapes (``e3b200.graphed``), the line
reports the cache hit rate of the timed region and the eager (no graph) time of the same batches.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload W2|W3|W4|W5]

``--workload`` selects the other BASELINE configs (lines committed under profiles/): W3 ``config_dipole`` (256
molecules), W4 ``config_diffusion`` score evaluation (128 molecules, complete graphs), W5 ``config_diffusion_CA`` (one
protein-sized graph of 2 000 residues, the neighbour list with the chain / random-pair criteria inside the step).

N > 1 is launched by the driver with torchrun (one rank per GPU); ranks hold different batches (graphs shard naturally,
no data-path collective; "weak" scaling).  The force-matching TRAINING step of W2 (flat-gradient NCCL all-reduce
overlapped with the backward, fused Adam) is measured at every N and reported under ``training``.
``--impl reference`` times the CPU oracle (the reference's dataflow restated; the genuine reference cannot be installed:
e3nn is not in the wheelhouse) on the host cores, on a bounded SAMPLE of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "equivariant-nn-zoo_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

ROTATE = 8                   # different batches per rank in the timed region
NODE_BUCKET, EDGE_BUCKET, MIN_PAD = 256, 4096, 128


class Workload:
    """one BASELINE config: model, synthetic batches, what a step is, the CPU sample"""
    name = config = metric = label = None
    graphs = 0
    cpu_graphs = 0
    seed = 0
    out_keys = ()
    pos_key = "pos"
    pre_edge = None
    graphable = False
    needs_grad = False

    def host_batch(self, seed):
        raise NotImplementedError

    def sample_batch(self):
        """bounded sample of the workload for the CPU arm"""
        raise NotImplementedError

    def sample_text(self, inputs, n_edges):
        raise NotImplementedError

    def meta(self):
        return {"config": self.config, "seed": self.seed}

    def build_model(self, dev):
        import product_harness
        return product_harness.build_product(self.meta(), torch.float32, dev)

    def attrs(self, tensors):
        import harness
        return harness.attrs_for(tensors)

    def units(self, tensors):
        return int(tensors[self.pos_key].shape[0])

    def eager_step(self, model, tensors):
        from e3_layers.data import Batch, computeEdgeIndex

        batch = Batch(self.attrs(tensors), **tensors)
        if self.pre_edge is not None:
            d, a = computeEdgeIndex(batch.data, batch.attrs, **self.pre_edge)
            batch.update(d)
            batch.attrs.update(a)
            batch = Batch(batch.attrs, **batch.data)
        if self.needs_grad:
            return model(batch)
        with torch.no_grad():
            return model(batch)

    def run_oracle(self, oracle, inputs):
        import harness
        ei = inputs.get("edge_index") if self.pre_edge is None else None
        inp = {k: v for k, v in inputs.items() if k != "edge_index"} if ei is not None else inputs
        if self.needs_grad:
            return harness.run_oracle(oracle, inp, torch.float32, pre_edge=self.pre_edge, edge_index=ei)
        with torch.no_grad():
            return harness.run_oracle(oracle, inp, torch.float32, pre_edge=self.pre_edge, edge_index=ei)


class W2(Workload):
    name, config, metric = "W2", "config_energy_force", "energy+force atoms/s"
    graphs, cpu_graphs, out_keys, pre_edge = 512, 32, ("energy", "forces"), {"r_max": 5.0}
    graphable, needs_grad = True, True
    label = ("W2: config_energy_force (n_dim 64, l_max 2, 5 interaction blocks, r_max 5.0) energy+force evaluation = "
             "neighbour list + forward + position-gradient backward, 512 synthetic QM9-shaped molecules per GPU")

    def host_batch(self, seed):
        from e3b200 import synthetic
        return synthetic.qm9_like(self.graphs, seed=seed)

    def sample_batch(self):
        from e3b200 import synthetic
        return synthetic.qm9_like(self.cpu_graphs, seed=0)

    def sample_text(self, inputs, n_edges):
        return (f"SAMPLE: {self.cpu_graphs} of the {self.graphs} {self.name} molecules ({inputs['pos'].shape[0]} atoms, "
                f"{n_edges} edges) per evaluation, neighbour list + forward + autograd forces, oracle (reference dataflow) fp32")


class W3(W2):
    name, config, metric = "W3", "config_dipole", "dipole-evaluation atoms/s"
    graphs, cpu_graphs, out_keys, seed = 256, 32, ("dipole",), 3
    needs_grad = False
    label = ("W3: config_dipole (l = 1 multipole head, odd-parity paths) evaluation = neighbour list + forward, "
             "256 synthetic molecules (17 species) per GPU")

    def host_batch(self, seed):
        from e3b200 import synthetic
        return synthetic.qm9_like(self.graphs, seed=seed, species_choices=tuple(range(1, 18)))

    def sample_batch(self):
        from e3b200 import synthetic
        return synthetic.qm9_like(self.cpu_graphs, seed=0, species_choices=tuple(range(1, 18)))

    def sample_text(self, inputs, n_edges):
        return (f"SAMPLE: {self.cpu_graphs} of the {self.graphs} W3 molecules ({inputs['pos'].shape[0]} atoms, {n_edges} edges) "
                f"per evaluation, neighbour list + forward, oracle (reference dataflow) fp32")


class W4(Workload):
    name, config, metric = "W4", "config_diffusion", "score-evaluation atoms/s"
    graphs, cpu_graphs, out_keys, seed = 128, 32, ("score",), 4
    graphable = True          # complete graphs come with the batch: one capture per batch shape, no host synchronisation
    label = ("W4: config_diffusion (VP-SDE score model, complete graphs, bond-type + time embeddings) batched score "
             "evaluation (no grad), 128 synthetic molecules per GPU")

    def meta(self):
        return {"config": self.config, "seed": self.seed, "spec": ""}

    def host_batch(self, seed):
        from e3b200 import synthetic
        return synthetic.diffusion_like(self.graphs, seed=seed)

    def sample_batch(self):
        from e3b200 import synthetic
        return synthetic.diffusion_like(self.cpu_graphs, seed=0)

    def sample_text(self, inputs, n_edges):
        return (f"SAMPLE: {self.cpu_graphs} of the {self.graphs} W4 molecules ({inputs['pos'].shape[0]} atoms, {n_edges} edges) "
                f"per score evaluation, oracle (reference dataflow) fp32")


class W5(Workload):
    name, config, metric = "W5", "config_diffusion_CA", "score-evaluation residues/s"
    graphs, out_keys, seed, pos_key = 1, ("score_CA",), 6, "CA"
    n_res, cpu_res = 2000, 250
    label = ("W5: config_diffusion_CA (8 blocks, LayerNormalization, avg 100 neighbours) score evaluation (no grad) of one "
             "protein-sized C-alpha graph of 2000 residues per GPU, INCLUDING the model's neighbour-list layer "
             "(radius OR same-chain |i-j| < 5 OR 2 % random pairs, evaluated in the kernel sweep)")

    def build_model(self, dev):
        from e3_layers import configs
        from e3_layers.utils import build
        from param_init import reseed_parameters

        model = build(configs.config_diffusion_CA().model_config)        # WITH the neighbour-list layer
        reseed_parameters(model, self.seed)
        return model.to(dev).eval()

    def host_batch(self, seed):
        from e3b200 import synthetic
        inp = synthetic.protein_like(self.n_res, seed=seed)
        inp.pop("edge_index"), inp.pop("_n_edges")
        return inp

    def sample_batch(self):
        from e3b200 import synthetic
        return synthetic.protein_like(self.cpu_res, seed=0)                 # with its seeded edge list

    def sample_text(self, inputs, n_edges):
        return (f"SAMPLE: one graph of {self.cpu_res} residues ({n_edges} edges; the workload has {self.n_res}) per score "
                f"evaluation, edge list given, oracle (reference dataflow) fp32")


WORKLOADS = {"W2": W2, "W3": W3, "W4": W4, "W5": W5}


def workload_config(wl, world, n_atoms, n_edges, extra=None):
    cfg = {"workload": wl.label, "graphs_per_gpu": wl.graphs, "atoms_per_gpu": n_atoms, "edges_per_gpu": n_edges,
           "batches": f"{ROTATE} different seeded batches per rank, rotated through the timed steps "
                      "(atoms / edges above: the first one)",
           "parallelism": f"graphs sharded over {world} rank(s), no data-path collective",
           "l2": "no explicit flush: per-layer per-edge weights are E*W*4 B = %s GB >> 126 MB L2, and consecutive steps "
                 "run different batches" % ("%.2f" % (n_edges * 1920 * 4 / 1e9) if n_edges is not None else "1.15")}
    if extra:
        cfg.update(extra)
    return cfg


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) == 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        sm = [float(r[0]) for r in self.rows]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def cpu_arm(wl, warmup, min_reps, budget_s, max_reps):
    """the oracle port of the reference's path on all host threads, on a bounded sample of the workload"""
    import harness

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inputs = wl.sample_batch()
    oracle = harness.build_oracle(wl.meta(), torch.float32)
    o = None
    for _ in range(max(1, warmup)):
        o = wl.run_oracle(oracle, inputs)
    reps, t0 = 0, time.perf_counter()
    while reps < min_reps or (time.perf_counter() - t0 < budget_s and reps < max_reps):
        o = wl.run_oracle(oracle, inputs)
        reps += 1
    dt = time.perf_counter() - t0
    n_edges = int(o["edge_index"].shape[1]) if "edge_index" in o else int(inputs["edge_index"].shape[1])
    value = wl.units(inputs) * reps / dt
    return {"value": value, "unit": wl.metric.split()[-1], "cores": cores, "kind": "port",
            "sample": wl.sample_text(inputs, n_edges) + f" x {reps} evaluations"}, dt / reps


def host_edge_count(wl, batch):
    """directed edges of a host batch without the GPU: per graph the ordered pairs i != j within r_max (reference
    data/compute_edge.py:56-75), or the edge list the batch brings; None if it cannot be told"""
    try:
        if wl.pre_edge is not None:
            pos, n_nodes = batch[wl.pos_key].double(), batch["_n_nodes"].reshape(-1).tolist()
            total, o = 0, 0
            for n in n_nodes:
                d = torch.cdist(pos[o:o + n], pos[o:o + n])
                total += int((d < wl.pre_edge["r_max"]).sum()) - n
                o += n
            return total
        if "edge_index" in batch:
            return int(batch["edge_index"].shape[1])
    except Exception:
        pass
    return None


def run_reference(args, wl):
    """CPU arm: the oracle port of the reference's path, all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu, s_per_step = cpu_arm(wl, args.warmup, args.steps, 0.0, args.steps)
    full = wl.host_batch(0)
    unit = wl.metric.split()[-1]
    full_edges = host_edge_count(wl, full)     # the `config` of both arms names the same workload
    line = {"impl": "reference", "metric": wl.metric, "value": cpu["value"], "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(wl, args.gpus, wl.units(full), full_edges,
                                      {"sample": "the CPU arm evaluates a bounded SAMPLE of this workload per step (see "
                                                 "cpu_baseline.sample); throughput is normalised per atom"}),
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def tp_bytes(E, N, st_mul_dims):
    W, D_in, D_mid = st_mul_dims
    return E * (4 * W + 4) + (N + 1) * 8 + N * (12 + 4 * D_in + 4 * D_mid)


def tp_bwd_bytes(E, N, st_mul_dims):
    """SURVEY 8d backward: per edge w read + dw written + Y read + dY written, per node x, dx, dmid + CSR"""
    W, D_in, D_mid = st_mul_dims
    return E * (8 * W + 8) + N * (24 + 8 * D_in + 4 * D_mid)


def training_step_time(model, batches, attrs, n_atoms_list, world, dev, dist, steps=None, warmup=None):
    """One optimiser step of config_energy_force per batch: neighbour list, forward, position gradient WITH
    its graph (second-order mode of GradientOutput), the reference's loss 1e3 MSE(E) + 3e4 MSE(F)
    (config_energy_force.py:30), backward to the parameters with the flat-gradient all-reduce issued bucket by bucket
    as the gradients become final (N > 1), Adam."""
    from e3_layers.data import Batch, computeEdgeIndex
    from e3b200 import optim

    # one untimed pass over every batch of the rotation first: the timed steps then run on shapes the caching allocator has
    # seen (with 2 warm-up steps the timed region paid cudaMalloc for new shapes on some hosts: 64 vs 96 ms per step)
    steps = steps or len(batches)
    warmup = warmup or len(batches)
    model.train()
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    opt = optim.FlatAdam(model, lr=1e-4)          # flat parameter / gradient buffers, fused Adam kernel
    if world > 1:
        opt.enable_overlap()
    g = torch.Generator().manual_seed(1)
    targets = []
    for b in batches:
        targets.append((torch.randn(b["_n_nodes"].shape[0], 1, generator=g).to(dev),
                        (0.1 * torch.randn(b["pos"].shape[0], 3, generator=g)).to(dev)))

    def step(i):
        resident, (e_t, f_t) = batches[i % len(batches)], targets[i % len(batches)]
        batch = Batch(dict(attrs), **{k: v.clone() for k, v in resident.items()})
        d, a = computeEdgeIndex(batch.data, batch.attrs, r_max=5.0)
        batch.update(d)
        batch.attrs.update(a)
        out = model(Batch(batch.attrs, **batch.data))
        loss = 1e3 * ((out["energy"] - e_t) ** 2).mean() + 3e4 * ((out["forces"] - f_t) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.all_reduce()
        opt.step()

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    atoms = torch.tensor([float(sum(n_atoms_list[(warmup + i) % len(batches)] for i in range(steps))) / steps], device=dev)
    # the collective on its own: the same flat buffer, back to back, device-timed
    ar_ms = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(atoms)
        for _ in range(3):
            dist.all_reduce(opt.grad)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            dist.all_reduce(opt.grad)
        a1.record()
        torch.cuda.synchronize()
        ar_ms = a0.elapsed_time(a1) / 10
    n_params = int(opt.grad.numel())
    model.load_state_dict(state)          # the optimiser steps above must not leak into anything measured later
    model.zero_grad(set_to_none=True)
    return {"what": "force-matching training step of the same workload over the same rotating batches: energy+force forward, "
                    "graph of the position gradient (second-order mode), loss 1e3 MSE(E) + 3e4 MSE(F), backward, flat-gradient "
                    "NCCL all-reduce issued per bucket from gradient hooks (overlaps the rest of the backward), fused Adam",
            "ms_per_step": float(t), "atoms_per_s": float(atoms) / (float(t) * 1e-3), "steps": steps, "warmup": warmup,
            "gradient_bytes": 4 * n_params,
            "nccl_allreduce_ms_standalone": ar_ms}


def run_ours(args, wl):
    import torch.distributed as dist

    from e3b200 import _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the ONE JSON line: whatever NCCL_DEBUG level the box sets (its version banner goes to stdout),
        # NCCL's own logging is sent to stderr
        # (NCCL honours NCCL_DEBUG_FILE only above the VERSION level, so VERSION / unset is raised to WARN)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    _lib.load()
    model = wl.build_model(dev)
    hosts = [wl.host_batch(rank * ROTATE + i) for i in range(ROTATE)]          # each rank its own batches
    attrs = wl.attrs(hosts[0])
    pinned = [{k: v.pin_memory() for k, v in h.items()} for h in hosts]
    n_atoms = [wl.units(h) for h in hosts]
    resident = [{k: v.to(dev) for k, v in h.items()} for h in hosts]

    from e3b200.graphed import GraphedEvaluator

    evaluator = None
    if wl.graphable and not args.eager:
        evaluator = GraphedEvaluator(model, r_max=wl.pre_edge["r_max"] if wl.pre_edge else 0.0, attrs=attrs, out_keys=wl.out_keys,
                                     node_bucket=NODE_BUCKET, edge_bucket=EDGE_BUCKET, min_pad_nodes=MIN_PAD,
                                     grad=wl.needs_grad, max_entries=ROTATE)

    def step(tensors, ev=True):
        if evaluator is not None and ev:              # public API: neighbour list eager, model step as a CUDA graph
            return evaluator(tensors)
        return wl.eager_step(model, tensors)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- set-up: every rotating batch once (plans, weight packs, graph captures -- the analogue of compilation) -----
    n_edges = []
    for r in resident:
        out = step({k: v.clone() for k, v in r.items()})
        if wl.pre_edge is not None:
            n_edges.append(int(ops.radius_graph(r[wl.pos_key], r["_n_nodes"].reshape(-1), wl.pre_edge["r_max"])[0].shape[1]))
        elif "edge_index" in r:
            n_edges.append(int(r["edge_index"].shape[1]))
        else:
            n_edges.append(int(out["edge_index"].shape[1]) if "edge_index" in out else 0)
    # ---- device-resident throughput ("value") -------------------------------------------------
    for i in range(args.warmup):
        out = step({k: v.clone() for k, v in resident[i % ROTATE].items()})
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:                 # one sampling thread per job: nvidia-smi processes on every rank would load the host
        sampler.start()
    hits0 = (evaluator.hits, evaluator.misses) if evaluator is not None else (0, 0)
    launches0 = _lib.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        out = step({k: v.clone() for k, v in resident[(args.warmup + i) % ROTATE].items()})
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count - launches0
    atoms_timed = float(sum(n_atoms[(args.warmup + i) % ROTATE] for i in range(args.steps)))
    hit_rate = None
    if evaluator is not None:
        h, m = evaluator.hits - hits0[0], evaluator.misses - hits0[1]
        hit_rate = h / max(1, h + m)
    # ---- the same batches op by op (no graph), with CUDA events around every kernel stage ---------------------------
    # (a replayed CUDA graph cannot be bracketed kernel by kernel: the per-kernel times of the roofline come from this
    #  eager pass of the same K steps, events on the launch stream)
    ops.TIMING = []
    eg0, eg1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eg0.record()
    for i in range(args.steps):
        step({k: v.clone() for k, v in resident[(args.warmup + i) % ROTATE].items()}, ev=False)
    eg1.record()
    sync_all()
    eager_ms = eg0.elapsed_time(eg1) / args.steps
    timing, ops.TIMING = ops.TIMING, None
    timing_how = ("CUDA events around every launch of the kernel in an eager pass of the same K steps run right after "
                  "the timed region (the timed region replays CUDA graphs of the step)" if evaluator is not None else
                  "CUDA events around every launch of the kernel (launch stream), eager pass of the same K steps")
    t = torch.tensor([ms, eager_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, eager_ms = float(t[0]), float(t[1])
    atoms_all = torch.tensor([atoms_timed], device=dev)
    if world > 1:
        dist.all_reduce(atoms_all)
    value = float(atoms_all.item()) / (ms * 1e-3)

    # ---- end to end through the public API with HOST buffers ("e2e") -----------------------------
    def e2e_step(i):
        dev_in = {k: v.to(dev, non_blocking=True) for k, v in pinned[i % ROTATE].items()}
        o = step(dev_in)
        return [o[k].cpu() for k in wl.out_keys]

    for i in range(max(1, args.warmup // 2)):
        e2e_step(i)
    sync_all()
    t0 = time.perf_counter()
    for i in range(args.steps):
        host_out = e2e_step(args.warmup + i)
    sync_all()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = float(atoms_all.item()) / float(t.item())
    clocks = sampler.stop() if rank == 0 else None      # sampled over the device-timed region, the eager kernel-timing pass and the e2e region
    h2d = sum(v.numel() * v.element_size() for v in pinned[0].values())
    d2h = sum(o.numel() * o.element_size() for o in host_out)

    # ---- the force-matching TRAINING step of the same workload (configs[1]), at every N ------------------------------
    training = None
    if wl.name == "W2" and not args.no_training:
        try:
            training = training_step_time(model, resident, attrs, n_atoms, world, dev, dist)
        except Exception as exc:                     # noqa: BLE001 -- reported in the line, never hides the headline
            if world > 1:
                raise
            training = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        model.eval()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: fused TP-conv forward of the heaviest layer structure ----------------------
    peak, peak_src = peaks()
    if args.breakdown:
        agg = {}
        for tag, s, e in timing:
            name = tag[1] if tag[0] == "stage" else ("f.tp_conv" if tag[0] == "fwd" else "b.tp_conv")
            agg[name] = agg.get(name, 0.0) + s.elapsed_time(e)
        tot = sum(agg.values())
        print("# per-step device time of the interaction blocks by stage (CUDA events on the launch stream), ms",
              file=sys.stderr)
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            print(f"#   {k:22s} {v / args.steps:8.3f}  {100 * v / tot:5.1f}%", file=sys.stderr)
        print(f"#   {'sum':22s} {tot / args.steps:8.3f}   (eager step {eager_ms:.3f} ms)", file=sys.stderr)

    def kernel_roofline(kind, bytes_fn):
        rows = [(s.elapsed_time(e), tag) for tag, s, e in timing if tag[0] == kind]
        if not rows:
            return None
        n_paths = max(tag[1] for _, tag in rows)                 # the heaviest structure of the model
        rows = [(ms_, tag) for ms_, tag in rows if tag[1] == n_paths]
        alg = sum(bytes_fn(tag[6], tag[5], (tag[1] * tag[2], tag[3], tag[4])) for _, tag in rows)
        tt = sum(ms_ for ms_, _ in rows)
        return rows, alg, tt

    roof = roof_bwd = None
    traffic_bwd = None
    f = kernel_roofline("fwd", tp_bytes)
    if f:
        rows, alg, tt = f
        _, n_paths, mul, x_dim, y_dim, N0, E0 = rows[0][1]
        ach = alg / (tt * 1e-3) / 1e9
        traffic = None     # dram bytes per launch of the same kernels from the committed ncu --set full capture
        for name in ("r2_tpfp_traffic.json",):
            tpath = os.path.join(ROOT, "profiles", name)
            if traffic is None and wl.name == "W2" and os.path.exists(tpath):
                with open(tpath) as fh:
                    tj = json.load(fh)
                if (tj.get("n_edges"), tj.get("n_nodes")) in {(tag[6], tag[5]) for _, tag in rows}:
                    traffic = tj["traffic_bytes_per_launch"]
                    bw = tj.get("backward_tpbp_S3")
                    if bw:
                        traffic_bwd = bw["dram_bytes_read"] + bw["dram_bytes_write"]
        all_fwd = sum(s.elapsed_time(e) for tag, s, e in timing if tag[0] == "fwd")
        roof = {"kernel": f"tpfp (fused gather + uvu CG tensor product + segmented sum, {n_paths} paths, mul {mul})",
                "note": "algorithmic bytes = SURVEY 8d per directed edge (one weight row per edge); since round 2 the two directions of "
                        "an undirected edge read ONE shared weight row (the second read mostly hits L2), so `traffic` (ncu DRAM bytes) "
                        "is below the algorithmic figure and `frac` can exceed what HBM alone would allow; the paired (FFMA2) kernel is "
                        "bound by the FMA pipe (ncu profiles/r2_ncu_tpfp2_S3.txt: packed instructions occupy it two cycles, ~77 % busy)",
                "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "peak_source": peak_src, "timing": timing_how,
                "algorithmic_bytes_per_launch": alg / len(rows), "avg_launch_ms": tt / len(rows),
                "launches_timed": len(rows), "edges_per_s": sum(tag[6] for _, tag in rows) / (tt * 1e-3),
                "share_of_step": all_fwd / ms if ms else None}
    b = kernel_roofline("bwd", tp_bwd_bytes)
    if b:
        rows, alg, tt = b
        seg = sum(s.elapsed_time(e) for tag, s, e in timing if tag[0] == "stage" and tag[1] == "b.segment_sum")
        n_seg = sum(1 for tag, s, e in timing if tag[0] == "stage" and tag[1] == "b.segment_sum")
        ach = alg / (tt * 1e-3) / 1e9
        all_bwd = sum(s.elapsed_time(e) for tag, s, e in timing if tag[0] == "bwd")
        roof_bwd = {"kernel": f"tpbp (backward of the same kernel: dw, dx, dY; {rows[0][1][1]} paths, mul {rows[0][1][2]})",
                    "note": "the single most expensive kernel of the step (share_of_step: all its launches / step time); "
                            "bound by instruction issue (ncu profiles/r2_ncu_tpbp1d_S2.txt: issue slots 72 % busy, FMA pipe 53 %, "
                            "top stall `not_selected`), not by HBM",
                    "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic_bwd if roof is not None and roof.get("traffic") else None,
                    "share_of_step": all_bwd / ms if ms else None,
                    "algorithmic_bytes_per_launch": alg / len(rows), "avg_launch_ms": tt / len(rows),
                    "launches_timed": len(rows),
                    "segment_sum_ms_per_step_all_layers": seg / args.steps if n_seg else 0.0}

    # ---- CPU baseline beside it (rank 0, N = 1 only): oracle on a bounded sample ------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_arm(wl, 1, 3, 10.0, 20)

    unit = wl.metric.split()[-1]
    execution = "eager (op by op)"
    if evaluator is not None and wl.pre_edge is None:
        execution = ("CUDA graph of the whole evaluation per input shape signature (the batch brings its edge list; no host "
                     "synchronisation; e3b200.graphed.GraphedEvaluator)")
    elif evaluator is not None:
        execution = ("CUDA graph of the model step per bucketed (atoms, edges, graphs) signature "
                     f"(node bucket {NODE_BUCKET}, edge bucket {EDGE_BUCKET}, >= {MIN_PAD} padding atoms; "
                     "e3b200.graphed.GraphedEvaluator), neighbour list eager")
    line = {"metric": wl.metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(wl, world, n_atoms[0], n_edges[0]),
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "execution": execution,
            "graph_cache": None if evaluator is None else {"hit_rate_timed_region": hit_rate, "captures": len(evaluator.cache),
                                                           "distinct_batches": ROTATE,
                                                           "distinct_shapes": len(set(zip(n_atoms, n_edges)))},
            "eager_ms_per_step": eager_ms,
            "clocks": clocks, "roofline": roof, "roofline_bwd": roof_bwd, "cpu_baseline": cpu, "training": training}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="W2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-training", dest="no_training", action="store_true",
                    help="skip the extra measurement of the force-matching training step")
    ap.add_argument("--training", action="store_true", help="(kept for compatibility: the training step is measured at every N)")
    ap.add_argument("--breakdown", action="store_true", help="print the per-stage device time table to stderr (implies --eager)")
    ap.add_argument("--eager", action="store_true", help="run the step op by op instead of replaying its CUDA graph")
    args = ap.parse_args()
    args.eager = args.eager or args.breakdown
    wl = WORKLOADS[args.workload]()
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        if args.gpus > 1 and "RANK" not in os.environ:
            # convenience: re-launch under torchrun when called directly with --gpus N
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        run_ours(args, wl)


if __name__ == "__main__":
    main()
