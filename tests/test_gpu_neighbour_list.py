"""Neighbour-list kernels on the B200 against the fixtures produced by the genuine reference
(bit-exact, same order) and against the oracle on larger seeded inputs."""
import json

import numpy as np
import pytest
import torch

import harness
from e3b200 import ops, synthetic
from oracle import ref_layers

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _check_csr(ei, csr, N):
    ei = ei.cpu()
    E = ei.shape[1]
    in_ptr, out_ptr = csr.in_ptr.cpu(), csr.out_ptr.cpu()
    in_eid = csr.in_eid.cpu().long() if csr.in_eid is not None else torch.arange(E)
    out_eid = csr.out_eid.cpu().long() if csr.out_eid is not None else torch.arange(E)
    assert sorted(in_eid.tolist()) == list(range(E)) and sorted(out_eid.tolist()) == list(range(E))
    dst_of_slot = torch.repeat_interleave(torch.arange(N), in_ptr[1:] - in_ptr[:-1])
    src_of_slot = torch.repeat_interleave(torch.arange(N), out_ptr[1:] - out_ptr[:-1])
    assert torch.equal(ei[1][in_eid], dst_of_slot)
    assert torch.equal(ei[0][out_eid], src_of_slot)
    assert torch.equal(csr.in_nbr.cpu().long(), ei[0][in_eid])


def test_radius_graph_matches_reference_fixtures():
    z = np.load(harness.GOLDEN + "/neighbour_lists.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    for name, r in meta.items():
        pos = torch.from_numpy(z[f"{name}/pos"]).to(DEV)
        n_nodes = torch.from_numpy(z[f"{name}/n_nodes"]).to(DEV)
        ei, n_edges, csr = ops.radius_graph(pos, n_nodes.reshape(-1), r)
        assert torch.equal(ei.cpu(), torch.from_numpy(z[f"{name}/edge_index"])), name   # bit-exact, same order
        assert torch.equal(n_edges.cpu(), torch.from_numpy(z[f"{name}/n_edges"])), name
        _check_csr(ei, csr, pos.shape[0])


@pytest.mark.parametrize("seed", [0, 1])
def test_radius_graph_vs_oracle_large(seed):
    b = synthetic.qm9_like(300, seed=seed)
    data = {"pos": b["pos"], "_n_nodes": b["_n_nodes"]}
    d, _ = ref_layers.computeEdgeIndex(data, {}, r_max=5.0)
    ei, n_edges, csr = ops.radius_graph(b["pos"].to(DEV), b["_n_nodes"].reshape(-1).to(DEV), 5.0)
    assert torch.equal(ei.cpu(), d["edge_index"])
    assert torch.equal(n_edges.cpu(), data["_n_edges"])
    p = synthetic.protein_like(1500, seed=seed)
    data = {"pos": p["CA"], "_n_nodes": p["_n_nodes"]}
    d, _ = ref_layers.computeEdgeIndex(data, {}, r_max=8.0 / 25.83)
    ei, _, csr = ops.radius_graph(p["CA"].to(DEV), p["_n_nodes"].reshape(-1).to(DEV), 8.0 / 25.83)
    assert torch.equal(ei.cpu(), d["edge_index"])
    _check_csr(ei, csr, 1500)


def test_empty_and_degenerate_graphs():
    pos = torch.zeros(0, 3, device=DEV)
    ei, n_edges, csr = ops.radius_graph(pos, torch.zeros(0, dtype=torch.long, device=DEV), 5.0)
    assert ei.shape == (2, 0)
    pos = torch.tensor([[0.0, 0, 0], [9.0, 0, 0]], device=DEV)      # no edges at all
    ei, n_edges, csr = ops.radius_graph(pos, torch.tensor([1, 1], device=DEV), 5.0)
    assert ei.shape == (2, 0) and n_edges.reshape(-1).tolist() == [0, 0]


def test_csr_of_arbitrary_edge_list():
    g = torch.Generator().manual_seed(3)
    N, E = 57, 900
    ei = torch.randint(0, N, (2, E), generator=g).to(DEV)
    csr = ops.build_csr(ei, N)
    _check_csr(ei, csr, N)
    # deterministic: ascending edge id inside every segment
    in_eid, in_ptr = csr.in_eid.cpu(), csr.in_ptr.cpu()
    for n in range(N):
        seg = in_eid[in_ptr[n]:in_ptr[n + 1]]
        assert torch.equal(seg, seg.sort().values)


def test_criteria_edges_bit_exact_vs_oracle():
    """computeEdgeIndex(criteria=...) (compute_edge.py:73-75; config_diffusion_CA.py:58-64 without its random part):
    radius edges OR same-chain |i - j| < 5 OR a seeded mask, self loops removed, reference order, _n_edges"""
    from e3_layers.data import computeEdgeIndex

    p = synthetic.protein_like(700, seed=3)
    two = {k: torch.cat([p[k], p[k]]) for k in ("CA", "chain_id")}            # two graphs in one batch
    two["chain_id"][700:] += 10
    n_nodes = torch.tensor([[700], [700]])
    mask_seed = torch.Generator().manual_seed(5)
    extra = torch.rand(2 * 700 * 700, generator=mask_seed) < 0.01                  # one entry per candidate pair

    def criteria(data, edge_index):
        src, dst = edge_index[0], edge_index[1]
        chain = data["chain_id"].view(-1)
        keep = (chain[src] == chain[dst]) & ((src - dst).abs() < 5)
        return keep | extra.to(src.device)

    ref_data = {"CA": two["CA"], "chain_id": two["chain_id"], "_n_nodes": n_nodes}
    d, _ = ref_layers.computeEdgeIndex(ref_data, {}, r_max=8.0 / 25.83, key="CA", criteria=criteria)
    data = {"CA": two["CA"].to(DEV), "chain_id": two["chain_id"].to(DEV), "_n_nodes": n_nodes.to(DEV)}
    out, attrs = computeEdgeIndex(data, {}, r_max=8.0 / 25.83, key="CA", criteria=criteria)
    assert torch.equal(out["edge_index"].cpu(), d["edge_index"])                   # same set, same order
    assert torch.equal(data["_n_edges"].cpu(), ref_data["_n_edges"])
    assert attrs["_n_edges"] == ("graph", "1x0e")
    csr = ops.graph_of(out["edge_index"], 1400)
    _check_csr(out["edge_index"], csr, 1400)


def test_pre_existing_edges_are_merged_and_attributes_remapped():
    """compute_edge.py:86-100 (the bonds of the dataset under the recomputed graph, config_diffusion.py:50 / :73-83):
    edge list and remapped, zero-padded per-edge tensors bit-exact against the oracle's restatement of the reference's
    double loop"""
    from e3_layers.data import computeEdgeIndex

    b = synthetic.qm9_like(9, seed=4, n_min=3, n_max=12)
    n = b["_n_nodes"].reshape(-1)
    g = torch.Generator().manual_seed(2)
    # a sparse symmetric "bond" list inside every molecule, in the reference's (source, destination) order
    src, dst, off = [], [], 0
    for k in n.tolist():
        a = torch.arange(off, off + k)
        i, j = torch.meshgrid(a, a, indexing="ij")
        keep = (torch.rand(k, k, generator=g) < 0.25)
        keep = (keep | keep.T) & (i != j)
        src.append(i[keep])
        dst.append(j[keep])
        off += k
    bonds = torch.stack([torch.cat(src), torch.cat(dst)])
    order = torch.argsort(bonds[0] * off + bonds[1])
    bonds = bonds[:, order]
    bond_type = torch.randint(1, 4, (bonds.shape[1], 1), generator=g)
    extra = torch.randn(bonds.shape[1], 3, generator=g)                     # a float per-edge tensor too
    attrs = {"pos": ("node", "1x1o"), "bond_type": ("edge", "1x0e"), "bond_vec": ("edge", "1x1o")}
    for r_max in (1.7, 9999.0):                                             # sparse radius graph / complete graph
        ref_data = {"pos": b["pos"], "_n_nodes": b["_n_nodes"], "edge_index": bonds.clone(), "bond_type": bond_type.clone(),
                    "bond_vec": extra.clone()}
        d, _ = ref_layers.computeEdgeIndex(ref_data, dict(attrs), r_max=r_max)
        data = {"pos": b["pos"].to(DEV), "_n_nodes": b["_n_nodes"].to(DEV), "edge_index": bonds.to(DEV),
                "bond_type": bond_type.to(DEV), "bond_vec": extra.to(DEV)}
        out, _ = computeEdgeIndex(data, dict(attrs), r_max=r_max)
        assert torch.equal(out["edge_index"].cpu(), d["edge_index"]), r_max
        assert torch.equal(data["bond_type"].cpu(), ref_data["bond_type"]) and data["bond_type"].dtype == torch.int64
        assert torch.equal(data["bond_vec"].cpu(), ref_data["bond_vec"])
        assert torch.equal(data["_n_edges"].cpu(), ref_data["_n_edges"])
        if r_max < 100:
            assert out["edge_index"].shape[1] > bonds.shape[1]              # the radius part really added edges


def test_pair_criteria_in_the_sweep_bit_exact_vs_oracle():
    """ops.PairCriteria (the criteria of config_diffusion_CA.py:58-64) evaluated inside the neighbour-list kernel
    (e3b_pair_graph_*): same edges, same order as the oracle's computeEdgeIndex fed the reference's own criteria
    function with the same uniforms; the foreign-callable path of the product agrees too."""
    from e3_layers.data import computeEdgeIndex

    p = synthetic.protein_like(600, seed=7)
    two = {k: torch.cat([p[k], p[k][:400]]) for k in ("CA", "chain_id")}           # two graphs, 600 + 400 residues
    two["chain_id"][600:] += 10
    n_nodes = torch.tensor([[600], [400]])
    u = torch.rand(600 * 600 + 400 * 400, generator=torch.Generator().manual_seed(9))

    def reference_criteria(data, edge_index):                                      # config_diffusion_CA.py:58-64, seeded
        mask = (data["chain_id"][edge_index[0]] == data["chain_id"][edge_index[1]]).view(-1)
        mask = torch.logical_and(mask, abs(edge_index[0] - edge_index[1]) < 5)
        return torch.logical_or(mask, u.to(mask.device) < 0.02)

    ref_data = {"CA": two["CA"], "chain_id": two["chain_id"], "_n_nodes": n_nodes}
    d, _ = ref_layers.computeEdgeIndex(ref_data, {}, r_max=8.0 / 25.83, key="CA", criteria=reference_criteria)
    crit = ops.PairCriteria(segment_key="chain_id", max_separation=5, p_random=0.02)
    data = {"CA": two["CA"].to(DEV), "chain_id": two["chain_id"].to(DEV), "_n_nodes": n_nodes.to(DEV), "_pair_uniforms": u.to(DEV)}
    out, _ = computeEdgeIndex(data, {}, r_max=8.0 / 25.83, key="CA", criteria=crit)
    assert torch.equal(out["edge_index"].cpu(), d["edge_index"])
    assert torch.equal(data["_n_edges"].cpu(), ref_data["_n_edges"])
    _check_csr(out["edge_index"], ops.graph_of(out["edge_index"], 1000), 1000)
    data2 = {k: v for k, v in data.items() if k != "_n_edges"}
    out2, _ = computeEdgeIndex(data2, {}, r_max=8.0 / 25.83, key="CA", criteria=lambda dd, ei: crit(dd, ei))
    assert torch.equal(out2["edge_index"], out["edge_index"])
    # without explicit uniforms: hash RNG seeded from torch's CPU generator -> repeatable, right density, no self loops
    data3 = {k: v for k, v in data.items() if k not in ("_n_edges", "_pair_uniforms")}
    torch.manual_seed(4)
    a, _ = computeEdgeIndex(dict(data3), {}, r_max=1e-6, key="CA", criteria=ops.PairCriteria(p_random=0.02))
    torch.manual_seed(4)
    b, _ = computeEdgeIndex(dict(data3), {}, r_max=1e-6, key="CA", criteria=ops.PairCriteria(p_random=0.02))
    c, _ = computeEdgeIndex(dict(data3), {}, r_max=1e-6, key="CA", criteria=ops.PairCriteria(p_random=0.02))
    assert torch.equal(a["edge_index"], b["edge_index"]) and not torch.equal(a["edge_index"], c["edge_index"])
    n_pairs = 600 * 599 + 400 * 399
    assert abs(a["edge_index"].shape[1] / n_pairs - 0.02) < 0.002
    assert bool((a["edge_index"][0] != a["edge_index"][1]).all())
    seg = torch.repeat_interleave(torch.arange(2, device=DEV), n_nodes.reshape(-1).to(DEV))
    assert torch.equal(seg[a["edge_index"][0]], seg[a["edge_index"][1]])           # never across graphs


@pytest.mark.parametrize("case", ["protein", "blob", "mixed"])
def test_cell_list_bit_exact_vs_all_pairs(case, monkeypatch):
    """the cell-list kernels (e3b_cell_graph_*) against the all-pairs sweep and the oracle: same edges, same order"""
    if case == "protein":
        p = synthetic.protein_like(3000, seed=2)
        pos, n_nodes, r = p["CA"], torch.tensor([3000]), 8.0 / 25.83
    elif case == "blob":                                      # dense: > 512 neighbours for many atoms (buffer overflow path)
        g = torch.Generator().manual_seed(3)
        pos, n_nodes, r = torch.rand(2500, 3, generator=g) * 4.0, torch.tensor([2500]), 1.9
    else:                                                     # a large graph, a flat one (one cell thick) and small ones
        g = torch.Generator().manual_seed(5)
        big = torch.rand(2000, 3, generator=g) * 30.0
        flat = torch.rand(1500, 3, generator=g) * torch.tensor([40.0, 40.0, 0.5])
        small = synthetic.qm9_like(20, seed=1)
        pos = torch.cat([big, small["pos"], flat, torch.zeros(1, 3)])
        n_nodes = torch.cat([torch.tensor([2000]), small["_n_nodes"].reshape(-1), torch.tensor([1500, 1])])
        r = 5.0
    monkeypatch.setattr(ops, "CELL_MIN_NODES", 10 ** 9)
    ei0, ne0, _ = ops.radius_graph(pos.to(DEV), n_nodes.to(DEV), r)
    monkeypatch.setattr(ops, "CELL_MIN_NODES", 1024 if case != "mixed" else 64)
    if case == "mixed":
        # the dispatch looks at the mean graph size; force the cell path for this mixed batch
        monkeypatch.setattr(ops, "_use_cells", lambda N, G: True)
    ei1, ne1, csr = ops.radius_graph(pos.to(DEV), n_nodes.to(DEV), r)
    assert torch.equal(ei1, ei0) and torch.equal(ne1, ne0)
    _check_csr(ei1, csr, pos.shape[0])
    if case == "protein":
        data = {"pos": pos, "_n_nodes": n_nodes.view(-1, 1)}
        d, _ = ref_layers.computeEdgeIndex(data, {}, r_max=r)
        assert torch.equal(ei1.cpu(), d["edge_index"])
    if case == "blob":
        deg = torch.bincount(ei1[0])
        assert int(deg.max()) > 512
