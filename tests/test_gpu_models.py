"""End-to-end parity of the product models on the B200 (through the C ABI) against the golden
fixtures of the genuine reference, against the oracle at larger sizes, and SO(3)/inversion
equivariance.  Tolerances: 1e-5 relative in fp32, 1e-10 in the fp64 mode (north_star)."""
import pytest
import torch

import harness
import product_harness
from e3b200 import synthetic
from oracle import wigner

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = [
    ("model_energy_force", ["energy", "forces"], {"r_max": 5.0}),
    ("model_energy", ["total_energy"], {"r_max": 4.0}),
    ("model_dipole", ["dipole"], {"r_max": 5.0}),
    ("model_diffusion", ["score"], None),
    ("model_diffusion_nll", ["score", "nll"], None),
    ("model_diffusion_CA", ["score_CA"], None),
]


@pytest.mark.parametrize("name,keys,pre_edge", CASES)
def test_models_fp64_vs_reference_fixtures(name, keys, pre_edge):
    g = harness.load_golden(name)
    model = product_harness.build_product(g["meta"], torch.float64, DEV)
    ei = g["out64"]["edge_index"] if name == "model_diffusion_CA" else None
    out = product_harness.run_product(model, g["in"], torch.float64, DEV, pre_edge=pre_edge, edge_index=ei)
    assert torch.equal(out["edge_index"].cpu(), g["out64"]["edge_index"])
    for k in keys + (["node_features"] if "node_features" in g["out64"] else []):
        err = harness.rel_err(out[k], g["out64"][k])
        assert err < (1e-8 if name == "model_diffusion_CA" else 1e-10), (name, k, err)   # D6, see test_composition_cpu


@pytest.mark.parametrize("name,keys,pre_edge", CASES)
def test_models_fp32_vs_reference_fixtures(name, keys, pre_edge):
    g = harness.load_golden(name)
    model = product_harness.build_product(g["meta"], torch.float32, DEV)
    ei = g["out32"]["edge_index"] if name == "model_diffusion_CA" else None
    out = product_harness.run_product(model, g["in"], torch.float32, DEV, pre_edge=pre_edge, edge_index=ei)
    assert torch.equal(out["edge_index"].cpu(), g["out32"]["edge_index"])
    if name == "model_diffusion_CA":
        # The fixture's fp32 and fp64 runs use different edge lists (fp32 vs fp64 cutoff predicate), so the
        # fp64 truth for THIS edge list is the product's own fp64 mode (held to the reference's fp64 run at 1e-8
        # above).  For this 8-block network the REFERENCE's fp32 run is itself 3.8e-5 away from that truth
        # (measured; fp32 evaluation of the inputs' embeddings, shared by every fp32 implementation), so the
        # bar is: no further from the fp64 truth than the reference's own fp32 run (+10 %), and within 1.5e-5
        # of that run (two fp32 evaluations; the op-by-op and fused product paths differ by 7e-6).
        model64 = product_harness.build_product(g["meta"], torch.float64, DEV)
        ref64 = product_harness.run_product(model64, g["in"], torch.float64, DEV, pre_edge=pre_edge, edge_index=ei)
        for k in keys:
            ours, theirs = harness.rel_err(out[k], ref64[k]), harness.rel_err(g["out32"][k], ref64[k])
            assert ours < max(1e-5, 1.1 * theirs), (name, k, ours, theirs)
            assert harness.rel_err(out[k], g["out32"][k]) < 1.5e-5, (name, k, "vs reference fp32 run")
        return
    for k in keys:
        # compare with the fp64 reference run: the fp32 reference itself carries ~1e-6 of noise
        err = harness.rel_err(out[k], g["out64"][k])
        assert err < 1e-5, (name, k, err)


def test_energy_force_vs_oracle_w1_batch():
    """a W1-sized batch (128 QM9-shaped molecules) at full width, fp32 product vs fp64 oracle"""
    meta = {"config": "config_energy_force", "seed": 3}
    inputs = synthetic.qm9_like(24, seed=5)
    oracle = harness.build_oracle(meta, torch.float64)
    ref = harness.run_oracle(oracle, inputs, torch.float64, pre_edge={"r_max": 5.0})
    model = product_harness.build_product(meta, torch.float32, DEV)
    out = product_harness.run_product(model, inputs, torch.float32, DEV, pre_edge={"r_max": 5.0})
    assert torch.equal(out["edge_index"].cpu(), ref["edge_index"])
    assert harness.rel_err(out["energy"], ref["energy"]) < 1e-5
    assert harness.rel_err(out["forces"], ref["forces"]) < 1e-5


@pytest.mark.parametrize("name,key", [("model_energy_force", "forces"), ("model_dipole", "dipole")])
def test_equivariance_on_gpu(name, key):
    g = harness.load_golden(name)
    model = product_harness.build_product(g["meta"], torch.float64, DEV)
    R = -wigner.rand_rotation(torch.Generator().manual_seed(9))        # rotation x inversion
    out0 = product_harness.run_product(model, g["in"], torch.float64, DEV, pre_edge={"r_max": 5.0})
    inp = dict(g["in"])
    inp["pos"] = inp["pos"].double() @ R.T
    out1 = product_harness.run_product(model, inp, torch.float64, DEV, pre_edge={"r_max": 5.0})
    assert harness.rel_err(out1[key].cpu(), out0[key].cpu() @ R.T) < 1e-10
    if "energy" in out0:
        assert harness.rel_err(out1["energy"], out0["energy"]) < 1e-12


def test_full_size_invariants_w2():
    """BASELINE W2 size (512 molecules): size-independent properties -- energy invariance under a
    rigid rotation, zero net force per molecule, permutation of molecules permutes energies."""
    meta = {"config": "config_energy_force", "seed": 4}
    model = product_harness.build_product(meta, torch.float32, DEV)
    inputs = synthetic.qm9_like(512, seed=0)
    out = product_harness.run_product(model, inputs, torch.float32, DEV, pre_edge={"r_max": 5.0})
    seg = out["_node_segment"].to(DEV)
    net = torch.zeros(512, 3, device=DEV).index_add_(0, seg, out["forces"])
    assert float(net.abs().max()) < 1e-4 * float(out["forces"].abs().max()) * 30
    R = wigner.rand_rotation(torch.Generator().manual_seed(2)).float()
    inp = dict(inputs)
    inp["pos"] = inputs["pos"] @ R.T
    out_r = product_harness.run_product(model, inp, torch.float32, DEV, pre_edge={"r_max": 5.0})
    assert harness.rel_err(out_r["energy"], out["energy"]) < 1e-5
    assert harness.rel_err(out_r["forces"].cpu(), out["forces"].cpu() @ R.T) < 1e-4


def test_graphed_evaluator_matches_eager():
    """CUDA-graph replay of the model step (e3b200.graphed) against the eager path: same shapes hit the cache,
    new inputs are honoured (different species, rigidly moved positions -> same edge count)."""
    from e3b200.graphed import GraphedEvaluator

    meta = {"config": "config_energy_force", "seed": 8}
    model = product_harness.build_product(meta, torch.float32, DEV)
    ev = GraphedEvaluator(model, r_max=5.0)
    base = synthetic.qm9_like(16, seed=6)
    variants = [base, dict(base), dict(base)]
    variants[1]["species"] = base["species"].flip(0).contiguous()
    variants[2]["pos"] = base["pos"] + torch.tensor([0.3, -0.2, 0.1])        # translation: same neighbour list
    for v in variants:
        eager = product_harness.run_product(model, v, torch.float32, DEV, pre_edge={"r_max": 5.0})
        got = ev({k: t.to(DEV) for k, t in v.items()})
        assert harness.rel_err(got["energy"], eager["energy"]) < 1e-6
        assert harness.rel_err(got["forces"], eager["forces"]) < 1e-5
    assert ev.misses == 1 and ev.hits == 2
    # a different shape misses and is captured separately
    other = synthetic.qm9_like(5, seed=1)
    got = ev({k: t.to(DEV) for k, t in other.items()})
    eager = product_harness.run_product(model, other, torch.float32, DEV, pre_edge={"r_max": 5.0})
    assert harness.rel_err(got["forces"], eager["forces"]) < 1e-5 and ev.misses == 2


def test_bucketed_evaluator_distinct_batches():
    """Bucketed padding (e3b200.graphed): DIFFERENT batches (different atom and edge counts) replay one captured graph;
    real atoms keep exactly their neighbours, so energies and forces equal the unpadded eager evaluation."""
    from e3b200.graphed import GraphedEvaluator

    meta = {"config": "config_energy_force", "seed": 8}
    model = product_harness.build_product(meta, torch.float32, DEV)
    ev = GraphedEvaluator(model, r_max=5.0, node_bucket=256, edge_bucket=4096, min_pad_nodes=16)
    shapes = set()
    for seed in (11, 12, 13, 14, 11):
        v = synthetic.qm9_like(24, seed=seed)
        eager = product_harness.run_product(model, v, torch.float32, DEV, pre_edge={"r_max": 5.0})
        shapes.add((eager["pos"].shape[0], eager["edge_index"].shape[1]))
        got = ev({k: t.to(DEV) for k, t in v.items()})
        assert got["energy"].shape == eager["energy"].shape and got["forces"].shape == eager["forces"].shape
        assert harness.rel_err(got["energy"], eager["energy"]) < 1e-6, seed
        assert harness.rel_err(got["forces"], eager["forces"]) < 1e-6, seed
    assert len(shapes) >= 4                       # the batches really differ ...
    assert ev.misses <= 2 and ev.hits >= 3        # ... and still share at most two captures
    # new weights: the stale captures are dropped and re-captured, never replayed
    with torch.no_grad():
        for p_ in model.parameters():
            p_.mul_(1.01)
    v = synthetic.qm9_like(24, seed=11)
    eager = product_harness.run_product(model, v, torch.float32, DEV, pre_edge={"r_max": 5.0})
    got = ev({k: t.to(DEV) for k, t in v.items()})
    assert harness.rel_err(got["forces"], eager["forces"]) < 1e-6


def test_shared_weight_rows_only_for_length_only_radial_embeddings():
    """the one-radial-MLP-per-undirected-edge path is taken when edge_radial provably depends on the edge length only
    (tag set by computeEdgeVector -> RadialBasisEncoding), equals the per-directed-edge path, and is NOT taken otherwise"""
    from e3b200 import interaction

    meta = {"config": "config_energy_force", "seed": 5}
    model = product_harness.build_product(meta, torch.float32, DEV)
    inputs = synthetic.qm9_like(12, seed=31)
    n0 = interaction.SHARED_CALLS
    shared = product_harness.run_product(model, inputs, torch.float32, DEV, pre_edge={"r_max": 5.0})
    n_blocks = interaction.SHARED_CALLS - n0
    assert n_blocks >= 4                                      # every full interaction block of the network
    interaction.SHARED_W = False
    try:
        n1 = interaction.SHARED_CALLS
        plain = product_harness.run_product(model, inputs, torch.float32, DEV, pre_edge={"r_max": 5.0})
        assert interaction.SHARED_CALLS == n1
    finally:
        interaction.SHARED_W = True
    assert torch.equal(shared["energy"], plain["energy"])     # bit-identical weights -> bit-identical forward
    assert harness.rel_err(shared["forces"], plain["forces"]) < 2e-6
    # the diffusion model's edge_radial is a concatenation with bond embeddings: no tag, no sharing
    meta = {"config": "config_diffusion", "seed": 4, "spec": ""}
    dm = product_harness.build_product(meta, torch.float32, DEV)
    d = synthetic.diffusion_like(4, seed=2)
    n2 = interaction.SHARED_CALLS
    with torch.no_grad():
        product_harness.run_product(dm, {k: v for k, v in d.items() if k != "edge_index"}, torch.float32, DEV, edge_index=d["edge_index"])
    assert interaction.SHARED_CALLS == n2


def test_graphed_evaluator_given_topology():
    """batches that bring their edge list (config_diffusion: complete graphs + bond types + t): the whole score evaluation is
    one captured graph per input shape, replays honour new inputs"""
    from e3b200.graphed import GraphedEvaluator

    meta = {"config": "config_diffusion", "seed": 4, "spec": ""}
    model = product_harness.build_product(meta, torch.float32, DEV)
    a = synthetic.diffusion_like(6, seed=1)
    ev = GraphedEvaluator(model, r_max=0.0, attrs=harness.attrs_for(a), out_keys=("score",), grad=False)
    b = dict(a)
    b["pos"] = a["pos"] * 0.9
    b["t"] = a["t"] * 0.5 + 0.1
    for v in (a, b, synthetic.diffusion_like(4, seed=2)):
        dev_in = {k: t.to(DEV) for k, t in v.items()}
        ei = dev_in.pop("edge_index")
        with torch.no_grad():
            eager = product_harness.run_product(model, {k: t for k, t in v.items() if k != "edge_index"}, torch.float32, DEV,
                                                edge_index=v["edge_index"])
        got = ev(dict(dev_in, edge_index=ei))
        assert harness.rel_err(got["score"], eager["score"]) < 1e-6
    assert ev.misses == 2 and ev.hits == 1


def test_padded_neighbour_list():
    """e3b_radius_graph_fill_padded: real atoms get exactly the edges of the plain radius graph, the padding rows are
    consistent (sorted edge list, row_ptr, reversed-edge index) and the total is the requested one"""
    from e3b200 import ops

    v = synthetic.qm9_like(9, seed=3)
    pos, nn = v["pos"].to(DEV), v["_n_nodes"].reshape(-1).to(DEV)
    ei0, _, _ = ops.radius_graph(pos, nn, 5.0)
    N = pos.shape[0]
    for n_pad, extra in ((8, 0), (8, 38), (7, 10), (2, 64)):
        ppos = torch.cat([pos, ops.padding_positions(n_pad, 5.0, DEV)])
        pnn = torch.cat([nn, torch.tensor([n_pad], device=DEV)])
        st = ops.radius_graph_count(ppos, pnn, 5.0)
        assert st.E == ei0.shape[1] + 2 * (n_pad // 2)
        Et = st.E + extra
        ei = torch.empty(2, Et, dtype=torch.int64, device=DEV)
        rev = torch.empty(Et, dtype=torch.int32, device=DEV)
        nbr = torch.empty(Et, dtype=torch.int32, device=DEV)
        ops.radius_graph_fill_padded(st, n_pad, Et, ei, rev, nbr)
        E0 = ei0.shape[1]
        assert torch.equal(ei[:, :E0], ei0)
        assert int(st.row_ptr[-1]) == Et and int(st.row_ptr[N]) == E0
        key = ei[0] * (N + n_pad) + ei[1]
        assert bool((key[1:] >= key[:-1]).all())                      # sorted (parallel edges are equal keys)
        assert bool((ei[:, E0:] >= N).all())
        deg = st.row_ptr[1:] - st.row_ptr[:-1]
        assert torch.equal(torch.bincount(ei[0], minlength=N + n_pad), deg)
        assert torch.equal(nbr.long(), ei[1])
        r = rev.long()
        assert torch.equal(ei[0][r], ei[1]) and torch.equal(ei[1][r], ei[0]) and torch.equal(r[r], torch.arange(Et, device=DEV))
        pd = deg[N:N + 2 * (n_pad // 2)]
        assert int(pd.max()) - int(pd.min()) <= 1                     # spread evenly over the pairs


def test_in_kernel_node_reduction_matches_deterministic_path():
    """evaluation-mode backward: d/dx reduced per source node by TMA reduce-adds inside the tensor-product kernel
    (e3b_tpconv_bwd_nodes) == per-edge rows + segment-sum kernel (E3B_DETERMINISTIC=1), to fp32 rounding"""
    from e3b200 import ops

    meta = {"config": "config_energy_force", "seed": 3}
    model = product_harness.build_product(meta, torch.float32, DEV)
    inputs = synthetic.qm9_like(40, seed=21)
    outs = {}
    for det in (True, False, False):
        ops.DETERMINISTIC = det
        try:
            o = product_harness.run_product(model, inputs, torch.float32, DEV, pre_edge={"r_max": 5.0})
        finally:
            ops.DETERMINISTIC = False
        outs.setdefault(det, []).append((o["energy"].clone(), o["forces"].clone()))
    (e_det, f_det), (e_a, f_a), (e_b, f_b) = outs[True][0], outs[False][0], outs[False][1]
    assert torch.equal(e_det, e_a)                                   # the forward pass is the same code
    assert harness.rel_err(f_a, f_det) < 2e-6 and harness.rel_err(f_b, f_a) < 2e-6
    oracle = harness.build_oracle(meta, torch.float64)
    ref = harness.run_oracle(oracle, inputs, torch.float64, pre_edge={"r_max": 5.0})
    assert harness.rel_err(f_a, ref["forces"]) < 1e-5


def test_ragged_batch_with_isolated_atoms():
    """single-atom molecules (no edges at all) and a far-apart pair inside an ordinary batch: empty CSR segments
    through every kernel, fp32 product vs fp64 oracle, forces of isolated atoms exactly zero"""
    meta = {"config": "config_energy_force", "seed": 6}
    inputs = synthetic.qm9_like(7, seed=8, n_min=1, n_max=9)
    n = inputs["_n_nodes"].reshape(-1).clone()
    n[0], n[3] = 1, 1                                     # force two single-atom graphs (re-partition the same atoms)
    n[-1] = inputs["pos"].shape[0] - int(n[:-1].sum())
    assert int(n[-1]) >= 1
    inputs["_n_nodes"] = n.view(-1, 1)
    inputs["pos"][int(n[0])] += 40.0                      # an atom of graph 1 far away from everything: no edges either
    oracle = harness.build_oracle(meta, torch.float64)
    ref = harness.run_oracle(oracle, inputs, torch.float64, pre_edge={"r_max": 5.0})
    model = product_harness.build_product(meta, torch.float32, DEV)
    out = product_harness.run_product(model, inputs, torch.float32, DEV, pre_edge={"r_max": 5.0})
    assert torch.equal(out["edge_index"].cpu(), ref["edge_index"])
    assert harness.rel_err(out["energy"], ref["energy"]) < 1e-5
    assert harness.rel_err(out["forces"], ref["forces"]) < 1e-5
    assert float(out["forces"][0].abs().max()) == 0.0 and float(out["forces"][int(n[0])].abs().max()) == 0.0
    # the same batch through the second-order (training) mode
    tr = product_harness.run_product(model.train(), inputs, torch.float32, DEV, pre_edge={"r_max": 5.0})
    assert harness.rel_err(tr["forces"], ref["forces"]) < 1e-5


def test_species_self_connection_matches_attribute_contraction():
    """node_attrs = PointwiseLinear(one-hot species) (reference embedCategorial, configs/layer_configs.py:8-30) are a
    function of the species: the self-connection then runs as one K = mul GEMM per path with the weights contracted per
    species (rows in species order, `ops.species_row_groups`) instead of the attribute-contraction epilogue.  Same
    numbers as the attribute path (fp32 rounding), in evaluation and in the second-order training mode, and not taken
    when the attributes are not tagged."""
    from e3b200 import interaction, ops

    meta = {"config": "config_energy_force", "seed": 9}
    model = product_harness.build_product(meta, torch.float32, DEV)
    inputs = synthetic.qm9_like(30, seed=13)
    n0 = interaction.SPECIES_SC_CALLS
    fast = product_harness.run_product(model, inputs, torch.float32, DEV, pre_edge={"r_max": 5.0})
    assert interaction.SPECIES_SC_CALLS - n0 >= 5              # every interaction block
    ops.SPECIES_SC = False
    try:
        n1 = interaction.SPECIES_SC_CALLS
        plain = product_harness.run_product(model, inputs, torch.float32, DEV, pre_edge={"r_max": 5.0})
        assert interaction.SPECIES_SC_CALLS == n1
        tr_plain = product_harness.run_product(model.train(), inputs, torch.float32, DEV, pre_edge={"r_max": 5.0})
    finally:
        ops.SPECIES_SC = True
    tr_fast = product_harness.run_product(model.train(), inputs, torch.float32, DEV, pre_edge={"r_max": 5.0})
    model.eval()
    assert harness.rel_err(fast["energy"], plain["energy"]) < 2e-6
    assert harness.rel_err(fast["forces"], plain["forces"]) < 2e-6
    assert harness.rel_err(tr_fast["forces"], tr_plain["forces"]) < 2e-6
    oracle = harness.build_oracle(meta, torch.float64)
    ref = harness.run_oracle(oracle, inputs, torch.float64, pre_edge={"r_max": 5.0})
    assert harness.rel_err(fast["energy"], ref["energy"]) < 1e-5
    assert harness.rel_err(fast["forces"], ref["forces"]) < 1e-5
    assert harness.rel_err(tr_fast["forces"], ref["forces"]) < 1e-5


def test_last_block_backward_restricted_to_scalar_paths():
    """An energy read-out consumes only the 0e scalars of the last interaction block: its linear map tags the gradient
    (`ops.tag_live_blocks`) and the block's backward runs the tensor-product plan restricted to the 3 paths that reach 0e
    (of 30), the matching column slice of the last radial layer and one path each of the post-reduction linear map and
    the self-connection.  Forces equal the unrestricted backward to fp32 rounding and the fp64 oracle to 1e-5; the
    restriction is not taken when something else also consumes the block's output."""
    from e3b200 import interaction

    meta = {"config": "config_energy_force", "seed": 11}
    model = product_harness.build_product(meta, torch.float32, DEV)
    inputs = synthetic.qm9_like(24, seed=17)
    n0 = interaction.SCALAR_ONLY_CALLS
    fast = product_harness.run_product(model, inputs, torch.float32, DEV, pre_edge={"r_max": 5.0})
    assert interaction.SCALAR_ONLY_CALLS - n0 == 2                # the last block (3 of 30 paths) and, through the tag it
                                                                  # hands on, the block before it (0e, 1o, 2e: 15 of 30)
    interaction.SCALAR_ONLY = False
    try:
        full = product_harness.run_product(model, inputs, torch.float32, DEV, pre_edge={"r_max": 5.0})
        assert interaction.SCALAR_ONLY_CALLS - n0 == 2
    finally:
        interaction.SCALAR_ONLY = True
    assert torch.equal(fast["energy"], full["energy"])
    assert harness.rel_err(fast["forces"], full["forces"]) < 2e-6
    oracle = harness.build_oracle(meta, torch.float64)
    ref = harness.run_oracle(oracle, inputs, torch.float64, pre_edge={"r_max": 5.0})
    assert harness.rel_err(fast["forces"], ref["forces"]) < 1e-5
    # a second consumer of the last block's features: autograd sums two gradients, the tag is gone, full backward
    from e3_layers.data import Batch, computeEdgeIndex
    data = harness.cast_inputs(inputs, torch.float32, DEV)
    batch = Batch(harness.attrs_for(data), **data)
    d, a = computeEdgeIndex(batch.data, batch.attrs, r_max=5.0)
    batch.update(d)
    batch.attrs.update(a)
    batch = Batch(batch.attrs, **batch.data)
    batch["pos"].requires_grad_(True)
    out = model.func(batch) if hasattr(model, "func") else None
    if out is not None:
        n1 = interaction.SCALAR_ONLY_CALLS
        y = out["energy"].sum() + 1e-3 * out["node_features"].pow(2).sum()
        from e3b200 import ops
        with ops.positions_only(batch["pos"]):
            (g,) = torch.autograd.grad(y, batch["pos"], retain_graph=True)
        assert interaction.SCALAR_ONLY_CALLS == n1 and bool(torch.isfinite(g).all())
        with ops.positions_only(batch["pos"]):                    # the read-out alone: restricted again
            (g2,) = torch.autograd.grad(out["energy"].sum(), batch["pos"])
        assert interaction.SCALAR_ONLY_CALLS == n1 + 2
        assert harness.rel_err(-g2, fast["forces"]) < 2e-6
