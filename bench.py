#!/usr/bin/env python3
"""Benchmark of the B200 hot path (BASELINE.json metric: energy+force atoms/s; fused TP-conv
edges/s and % of the HBM roofline).

A "step" = one energy+force evaluation of the reference's ``config_energy_force`` model
(n_dim 64, l_max 2, 5 interaction blocks, r_max 5) on one batch of 512 synthetic QM9-shaped
molecules per GPU (workload W2): neighbour list + forward + position-gradient
backward, through the public ``e3_layers`` API (``GradientOutput.forward``).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by the driver with torchrun (one rank per GPU); ranks hold different batches
(graphs shard naturally, no data-path collective; "weak" scaling).  ``--impl reference`` times
the CPU oracle (the reference's dataflow restated; the genuine reference cannot be installed:
e3nn is not in the wheelhouse) on the host cores, on a bounded sample of the same workload.
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

GRAPHS_PER_GPU = 512
W2_EDGES = 149452            # edges of the seed-0 W2 batch at r_max 5 (what the CUDA arm reports as edges_per_gpu)
CPU_SAMPLE_GRAPHS = 32
METRIC = "energy+force atoms/s"
META = {"config": "config_energy_force", "seed": 0}


WORKLOAD = ("W2: config_energy_force (n_dim 64, l_max 2, 5 interaction blocks, r_max 5.0) energy+force evaluation = "
            "neighbour list + forward + position-gradient backward, 512 synthetic QM9-shaped molecules per GPU")


def workload_config(world, n_atoms, n_edges):
    return {"workload": WORKLOAD, "graphs_per_gpu": GRAPHS_PER_GPU, "atoms_per_gpu": n_atoms, "edges_per_gpu": n_edges,
            "parallelism": f"graphs sharded over {world} rank(s), no data-path collective",
            "l2": "no explicit flush: per-layer per-edge weights are E*W*4 B = %.2f GB >> 126 MB L2" % (n_edges * 1920 * 4 / 1e9)}


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
def run_reference(args):
    """CPU arm: the oracle port of the reference's path, all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import harness
    from e3b200 import synthetic

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inputs = synthetic.qm9_like(CPU_SAMPLE_GRAPHS, seed=0)
    model = harness.build_oracle(META, torch.float32)
    n_atoms = inputs["pos"].shape[0]
    n_edges = None
    for _ in range(args.warmup):
        out = harness.run_oracle(model, inputs, torch.float32, pre_edge={"r_max": 5.0})
        n_edges = out["edge_index"].shape[1]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = harness.run_oracle(model, inputs, torch.float32, pre_edge={"r_max": 5.0})
        n_edges = out["edge_index"].shape[1]
    dt = time.perf_counter() - t0
    value = n_atoms * args.steps / dt
    sample = (f"{CPU_SAMPLE_GRAPHS} of the {GRAPHS_PER_GPU} W2 molecules ({n_atoms} atoms, {n_edges} edges) per step, "
              f"neighbour list + forward + autograd forces, fp32")
    full = synthetic.qm9_like(GRAPHS_PER_GPU, seed=0)            # the arm's workload (rank 0 batch), for the config block
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "atoms/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, full["pos"].shape[0], W2_EDGES),
            "cpu_baseline": {"value": value, "unit": "atoms/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "atoms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def tp_bytes(E, N, st_mul_dims):
    W, D_in, D_mid = st_mul_dims
    return E * (4 * W + 4) + (N + 1) * 8 + N * (12 + 4 * D_in + 4 * D_mid)


def training_step_time(model, resident, attrs, Batch, computeEdgeIndex, n_atoms, world, dev, dist, steps=4, warmup=2):
    """One optimiser step of config_energy_force on the same batch: neighbour list, forward, position gradient WITH
    its graph (second-order mode of GradientOutput), the reference's loss 1e3 MSE(E) + 3e4 MSE(F)
    (config_energy_force.py:30), backward to the parameters, flat-gradient all-reduce (N > 1), Adam."""
    from e3b200 import optim

    model.train()
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    opt = optim.FlatAdam(model, lr=1e-4)          # flat parameter / gradient buffers, fused Adam kernel
    g = torch.Generator().manual_seed(1)
    e_t = torch.randn(resident["_n_nodes"].shape[0], 1, generator=g).to(dev)
    f_t = (0.1 * torch.randn(n_atoms, 3, generator=g)).to(dev)

    def step():
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

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    atoms = torch.tensor([float(n_atoms)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(atoms)
    model.load_state_dict(state)          # the optimiser steps above must not leak into anything measured later
    model.zero_grad(set_to_none=True)
    return {"what": "force-matching training step of the same workload: energy+force forward, graph of the position gradient "
                    "(second-order mode), loss 1e3 MSE(E) + 3e4 MSE(F), backward, gradient all-reduce, Adam",
            "ms_per_step": float(t), "atoms_per_s": float(atoms) / (float(t) * 1e-3), "steps": steps, "warmup": warmup}


def run_ours(args):
    import torch.distributed as dist

    import product_harness
    from e3_layers.data import Batch, computeEdgeIndex
    from e3b200 import _lib, ops, synthetic

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
    model = product_harness.build_product(META, torch.float32, dev)
    host = synthetic.qm9_like(GRAPHS_PER_GPU, seed=rank)       # each rank its own molecules
    attrs = {"pos": ("node", "1x1o"), "species": ("node", "1x0e"), "_n_nodes": ("graph", "1x0e")}
    pinned = {k: v.pin_memory() for k, v in host.items()}
    n_atoms = host["pos"].shape[0]
    resident = {k: v.to(dev) for k, v in host.items()}

    from e3b200.graphed import GraphedEvaluator

    evaluator = GraphedEvaluator(model, r_max=5.0, attrs=attrs) if not args.eager else None

    def step(tensors):
        if evaluator is not None:                     # public API: neighbour list eager, model step as a CUDA graph
            return evaluator(tensors)
        batch = Batch(dict(attrs), **{k: v for k, v in tensors.items()})
        d, a = computeEdgeIndex(batch.data, batch.attrs, r_max=5.0)
        batch.update(d)
        batch.attrs.update(a)
        batch = Batch(batch.attrs, **batch.data)
        return model(batch)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput ("value") -------------------------------------------------
    for _ in range(args.warmup):
        out = step({k: v.clone() for k, v in resident.items()})
    n_edges = int(ops.radius_graph(resident["pos"], resident["_n_nodes"].reshape(-1), 5.0)[0].shape[1])
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:                 # one sampling thread per job: nvidia-smi processes on every rank would load the host
        sampler.start()
    ops.TIMING = []            # (tag, start_event, end_event) per fused TP-conv launch
    launches0 = _lib.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = step({k: v.clone() for k, v in resident.items()})
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count - launches0
    timing, ops.TIMING = ops.TIMING, None
    timing_how = "CUDA events around every launch of the kernel inside the timed region (launch stream)"
    if evaluator is not None:
        # a replayed CUDA graph cannot be bracketed kernel by kernel: time the SAME kernels on the SAME inputs in
        # an eager pass of the same K steps right after the timed region (events on the launch stream)
        ops.TIMING = []
        evaluator, keep = None, evaluator
        for _ in range(args.steps):
            step({k: v.clone() for k, v in resident.items()})
        sync_all()
        evaluator = keep
        timing, ops.TIMING = ops.TIMING, None
        timing_how = ("CUDA events around every launch of the kernel in an eager pass of the same K steps run right after "
                      "the timed region (the timed region replays a CUDA graph of the step)")
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    atoms_all = torch.tensor([float(n_atoms)], device=dev)
    if world > 1:
        dist.all_reduce(atoms_all)
    value = float(atoms_all.item()) * args.steps / (ms * 1e-3)

    # ---- end to end through the public API with HOST buffers ("e2e") -----------------------------
    def e2e_step():
        dev_in = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        o = step(dev_in)
        return o["energy"].cpu(), o["forces"].cpu()

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e_host, f_host = e2e_step()
    sync_all()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = float(atoms_all.item()) * args.steps / float(t.item())
    clocks = sampler.stop() if rank == 0 else None      # sampled over the device-timed region, the eager kernel-timing pass and the e2e region
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    d2h = e_host.numel() * e_host.element_size() + f_host.numel() * f_host.element_size()

    # ---- the force-matching TRAINING step of the same workload (configs[1]), reported beside the headline ----------
    # (by default only at N = 1, like the CPU baseline: the headline line of a multi-rank run must not depend on an
    #  extra with its own collectives; --training forces it, profiles/r1_bench_v6_2gpu.json was taken that way)
    training = None
    if not args.no_training and (world == 1 or args.training):
        try:
            training = training_step_time(model, resident, attrs, Batch, computeEdgeIndex, n_atoms, world, dev, dist)
        except Exception as exc:                     # noqa: BLE001 -- reported in the line, never hides the headline
            if world > 1:
                raise
            training = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        model.eval()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: fused TP-conv forward of a full (30-path) layer ----------
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
        print(f"#   {'sum':22s} {tot / args.steps:8.3f}   (step {ms / args.steps:.3f} ms)", file=sys.stderr)
    full = [(s.elapsed_time(e), tag) for tag, s, e in timing if tag[0] == "fwd" and tag[1] == 30]
    bwd = [(s.elapsed_time(e), tag) for tag, s, e in timing if tag[0] == "bwd" and tag[1] == 30]
    roof = None
    if full:
        t_ms = statistics.mean(x[0] for x in full)
        _, n_paths, mul, x_dim, y_dim, N, E = full[0][1]
        alg = tp_bytes(E, N, (n_paths * mul, x_dim, y_dim))
        ach = alg / (t_ms * 1e-3) / 1e9
        traffic = None     # dram bytes per launch of the same kernel from the committed ncu --set full capture
        tpath = os.path.join(ROOT, "profiles", "r1_tpfp_S3_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("n_edges") == E and tj.get("n_nodes") == N:
                traffic = tj["traffic_bytes_per_launch"]
        roof = {"kernel": "tpfp_S3<64> (fused gather + uvu CG tensor product + segmented sum, 30 paths, mul 64)",
                "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "peak_source": peak_src, "timing": timing_how, "algorithmic_bytes_per_launch": alg, "avg_launch_ms": t_ms,
                "launches_timed": len(full), "edges_per_s": E / (t_ms * 1e-3),
                "bwd_avg_launch_ms": statistics.mean(x[0] for x in bwd) if bwd else None}

    # ---- CPU baseline beside it (rank 0, N = 1 only): oracle on a bounded sample ------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import harness

        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sample_in = synthetic.qm9_like(CPU_SAMPLE_GRAPHS, seed=0)
        oracle = harness.build_oracle(META, torch.float32)
        harness.run_oracle(oracle, sample_in, torch.float32, pre_edge={"r_max": 5.0})   # warm-up
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 10 and reps < 20):
            o = harness.run_oracle(oracle, sample_in, torch.float32, pre_edge={"r_max": 5.0})
            reps += 1
        dt = time.perf_counter() - t0
        cpu = {"value": sample_in["pos"].shape[0] * reps / dt, "unit": "atoms/s", "cores": cores, "kind": "port",
               "sample": f"{CPU_SAMPLE_GRAPHS} of the {GRAPHS_PER_GPU} W2 molecules ({sample_in['pos'].shape[0]} atoms, "
                         f"{o['edge_index'].shape[1]} edges) x {reps} evaluations, oracle (reference dataflow) fp32"}

    line = {"metric": METRIC, "value": value, "unit": "atoms/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, n_atoms, n_edges),
            "e2e": {"value": e2e_value, "unit": "atoms/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "execution": "eager" if args.eager else "CUDA graph of the model step per (atoms, "
            "edges, graphs) signature, neighbour list eager (e3b200.graphed.GraphedEvaluator)",
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "training": training}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-training", dest="no_training", action="store_true",
                    help="skip the extra measurement of the force-matching training step")
    ap.add_argument("--training", action="store_true", help="measure the training step at N > 1 too")
    ap.add_argument("--breakdown", action="store_true", help="print the per-stage device time table to stderr (implies --eager)")
    ap.add_argument("--eager", action="store_true", help="run the step op by op instead of replaying its CUDA graph")
    args = ap.parse_args()
    args.eager = args.eager or args.breakdown
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.gpus > 1 and "RANK" not in os.environ:
            # convenience: re-launch under torchrun when called directly with --gpus N
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        run_ours(args)


if __name__ == "__main__":
    main()
