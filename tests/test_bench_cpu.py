"""Host-side pieces of bench.py that run without a GPU: the `config` object of the reference (CPU) arm -- it once crashed on a
missing edge count, which would have left the driver without a reference line -- and the host edge count behind it."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import bench  # noqa: E402


def test_workload_config_without_edge_count():
    for wl_cls in bench.WORKLOADS.values():
        cfg = bench.workload_config(wl_cls(), 1, 100, None, {"sample": "x"})
        assert cfg["edges_per_gpu"] is None and "workload" in cfg and cfg["sample"] == "x"


def test_host_edge_count_matches_the_oracle_neighbour_list():
    from oracle import ref_layers
    wl = bench.WORKLOADS["W2"]()
    batch = wl.sample_batch()
    n = bench.host_edge_count(wl, batch)
    data = {"pos": batch["pos"].double(), "_n_nodes": batch["_n_nodes"]}
    d, _ = ref_layers.computeEdgeIndex(dict(data), {"pos": ("node", "1x1o")}, r_max=wl.pre_edge["r_max"])
    assert n == int(d["edge_index"].shape[1]) and n > 0


def test_host_edge_count_of_a_batch_with_its_edge_list():
    wl = bench.WORKLOADS["W5"]()
    batch = wl.sample_batch()
    assert bench.host_edge_count(wl, batch) == int(batch["edge_index"].shape[1])
