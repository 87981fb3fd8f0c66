"""Dataset -> device pipeline (reference e3_layers/data/dataset.py, dataloader.py): host logic on the
CPU (index arithmetic, split, rank shard, file round trip, the preprocess contract) and, on the GPU, bit-for-bit
equality with the direct ``Batch`` path including a training run fed from the pipeline."""
import os
import sys
from functools import partial

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from e3_layers.data import Batch, CondensedDataset, DevicePipeline, getDataIters
from e3_layers.data.dataloader import rank_paths, split_indices
from e3_layers.data.dataset import save_npz
from e3b200 import synthetic

ATTRS_MOL = {"pos": ("node", "1x1o"), "species": ("node", "1x0e"), "_n_nodes": ("graph", "1x0e")}
ATTRS_DIFF = dict(ATTRS_MOL, t=("graph", "1x0e"), bond_type=("edge", "1x0e"), _n_edges=("graph", "1x0e"))


def _direct(ds, graphs, attrs):
    return Batch.from_data_list([ds.get(int(i)) for i in graphs], dict(attrs))


@pytest.mark.parametrize("resident", [True, False])
def test_batches_equal_collated_items_cpu(resident):
    """vectorised gathers (incl. re-based edge_index and per-edge tensors) == collating the individual items"""
    ds = CondensedDataset(data=synthetic.diffusion_like(23, seed=1), attrs=dict(ATTRS_DIFF))
    mk = lambda: DevicePipeline(ds, batch_size=3, shuffle=True, generator=torch.Generator().manual_seed(5), device="cpu",
                                resident=resident, world_size=2, rank=1, drop_last=False)
    idx = list(mk().batches_of_indices())
    assert len(idx) == 4 and sum(len(g) for g in idx) == 11          # rank 1 of 2 takes every second graph
    n = 0
    for b, g in zip(mk(), idx):
        ref = _direct(ds, g, ATTRS_DIFF)
        assert set(ref.keys()) == set(b.keys())
        for k in ref.keys():
            assert torch.equal(ref[k], b[k]), k
        n += 1
    assert n == 4


def test_ranks_partition_every_global_batch():
    ds = CondensedDataset(data=synthetic.qm9_like(40, seed=2), attrs=dict(ATTRS_MOL))
    seen = []
    for r in range(4):
        p = DevicePipeline(ds, batch_size=2, shuffle=True, generator=torch.Generator().manual_seed(9), device="cpu",
                           world_size=4, rank=r, drop_last=True)
        seen.append([g.tolist() for g in p.batches_of_indices()])
    assert all(len(s) == 5 for s in seen)
    for b in range(5):
        union = sorted(sum((seen[r][b] for r in range(4)), []))
        assert len(union) == 8 and len(set(union)) == 8
    assert sorted(sum((sum(s, []) for s in seen), [])) == list(range(40))


def test_endless_skip_replays_the_order():
    ds = CondensedDataset(data=synthetic.qm9_like(12, seed=3), attrs=dict(ATTRS_MOL))
    mk = lambda: DevicePipeline(ds, batch_size=5, shuffle=True, generator=torch.Generator().manual_seed(1), device="cpu",
                                drop_last=True)
    it = mk().endless()
    full = [next(it)["pos"] for _ in range(7)]                     # 2 batches per epoch -> spans epochs
    it2 = mk().endless(skip=3)
    for want in full[3:]:
        assert torch.equal(next(it2)["pos"], want)


def test_endless_refuses_an_empty_epoch():
    ds = CondensedDataset(data=synthetic.qm9_like(3, seed=3), attrs=dict(ATTRS_MOL))
    p = DevicePipeline(ds, batch_size=8, drop_last=True, device="cpu")
    assert len(p) == 0 and list(p) == []
    with pytest.raises(ValueError):
        next(p.endless())


def test_preprocess_contract_is_merged_not_replaced():
    """a layer-style function returning only its NEW tensors keeps the item's other tensors
    (the reference loses them: dataset.py:115-117 with compute_edge.py:110-113)"""
    def fake_edges(data, attrs):
        n = data["pos"].shape[0]
        attrs = dict(attrs, _n_edges=("graph", "1x0e"))
        return {"edge_index": torch.zeros(2, n, dtype=torch.long), "_n_edges": torch.tensor([[n]])}, attrs

    ds = CondensedDataset(data=synthetic.qm9_like(5, seed=4), attrs=dict(ATTRS_MOL), preprocess=[fake_edges])
    item = ds[2]
    assert {"pos", "species", "edge_index", "_n_edges"} <= set(item.keys())
    assert item["edge_index"].shape[1] == item["pos"].shape[0]
    one_arg = CondensedDataset(data=synthetic.qm9_like(5, seed=4), attrs=dict(ATTRS_MOL), preprocess=[lambda d: d])
    assert torch.equal(one_arg[1]["pos"], ds.get(1)["pos"])


def test_npz_round_trip_directory_and_regexp(tmp_path):
    a = CondensedDataset(data=synthetic.qm9_like(4, seed=5), attrs=dict(ATTRS_MOL))
    b = CondensedDataset(data=synthetic.qm9_like(3, seed=6), attrs=dict(ATTRS_MOL))
    save_npz(str(tmp_path / "part0.npz"), a)
    save_npz(str(tmp_path / "part1.npz"), b)
    save_npz(str(tmp_path / "other.npz"), b)
    one = CondensedDataset(path=str(tmp_path / "part0.npz"))
    assert torch.equal(one["pos"], a["pos"]) and one.attrs["pos"] == ("node", "1x1o")
    both = CondensedDataset(path=str(tmp_path) + ":.*part\\d\\.npz")
    assert len(both) == 7 and int(both["_n_nodes"].sum()) == both["pos"].shape[0]
    assert torch.equal(both["pos"], torch.cat([a["pos"], b["pos"]]))
    mapped = CondensedDataset(path=str(tmp_path / "part0.npz"), key_map={"species": "Z"})
    assert "Z" in mapped.keys() and "species" not in mapped.keys() and mapped.attrs["Z"] == ("node", "1x0e")
    lst = CondensedDataset(path=[str(tmp_path / "part0.npz"), str(tmp_path / "part1.npz")])
    assert len(lst) == 7


def test_split_and_rank_paths():
    tr, va = split_indices(100, 0.8, 10, "sequential")
    assert tr.tolist() == list(range(80)) and va.tolist() == list(range(80, 90))
    tr, va = split_indices(50, 30, 20, "random", torch.Generator().manual_seed(0))
    assert sorted(tr.tolist() + va.tolist()) == list(range(50))
    with pytest.raises(ValueError):
        split_indices(10, 8, 8, "random")
    paths = [f"p{i}" for i in range(8)]
    assert rank_paths(paths, 3, 8) == ["p3"] and rank_paths(paths, 1, 4) == ["p2", "p3"] and rank_paths("x", 0, 4) == "x"


def test_statistics_modes():
    d = synthetic.qm9_like(30, seed=7)
    n = d["_n_nodes"].reshape(-1)
    seg = torch.repeat_interleave(torch.arange(30), n)
    per_species = torch.arange(0, 120, dtype=torch.float32) * 0.5
    d["energy"] = torch.zeros(30).index_add_(0, seg, per_species[d["species"].reshape(-1)]).view(-1, 1)
    ds = CondensedDataset(data=d, attrs=dict(ATTRS_MOL, energy=("graph", "1x0e")))
    (uniq, counts), (rms,), (mean, std) = ds.statistics(["species-count", "pos-rms", "energy-per-species-mean_std"])
    assert int(counts.sum()) == d["pos"].shape[0]
    assert abs(float(rms) - float(d["pos"].pow(2).mean().sqrt())) < 1e-6
    for z in uniq.tolist():                                    # the composition regression recovers the per-species energy
        assert abs(float(mean[z, 0]) - 0.5 * z) < 1e-3
    assert float(std.max()) < 1e-3


def test_get_data_iters_cpu():
    class C:
        pass
    cfg, cfg.batch_size, cfg.data_config = C(), 4, C()
    cfg.data_config.n_train, cfg.data_config.n_val, cfg.data_config.train_val_split = 0.75, 0.25, "sequential"
    ds = CondensedDataset(data=synthetic.qm9_like(16, seed=8), attrs=dict(ATTRS_MOL))
    tr, va = getDataIters(cfg, seed=3, device="cpu", dataset=ds)
    b = next(va)
    assert b["_n_nodes"].numel() == 4 and torch.equal(b["pos"], ds.index_select([12, 13, 14, 15])["pos"])
    got = sorted(int(x) for _ in range(3) for x in next(tr)["_n_nodes"].reshape(-1))
    assert got == sorted(ds["_n_nodes"].reshape(-1)[:12].tolist())


# ---------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("resident", [True, False])
def test_pipeline_batches_on_device_bit_exact(resident):
    """device batches (HBM-resident dataset / pinned + copy stream) with the neighbour list as a preprocess ==
    collating the items on the host, moving them and calling computeEdgeIndex"""
    from e3_layers.data import computeEdgeIndex

    ds = CondensedDataset(data=synthetic.qm9_like(60, seed=9), attrs=dict(ATTRS_MOL),
                          preprocess=[partial(computeEdgeIndex, r_max=5.0)])
    mk = lambda: DevicePipeline(ds, batch_size=16, shuffle=True, generator=torch.Generator().manual_seed(2), device="cuda",
                                resident=resident, drop_last=True)
    idx = list(mk().batches_of_indices())
    for b, g in zip(mk(), idx):
        ref = _direct(ds, g, ATTRS_MOL).to("cuda")
        d, a = computeEdgeIndex(ref.data, ref.attrs, r_max=5.0)
        assert b["pos"].is_cuda and torch.equal(b["pos"], ref["pos"]) and torch.equal(b["species"], ref["species"])
        assert torch.equal(b["edge_index"], d["edge_index"]) and torch.equal(b["_n_edges"], ref["_n_edges"])


@pytest.mark.gpu
def test_training_fed_from_the_pipeline_matches_direct_batches():
    """train.py's loop fed by the pipeline == the same loop fed by directly built batches, bit for bit"""
    import train
    from e3_layers import configs
    from e3_layers.data import computeEdgeIndex
    from e3_layers.utils import build, setSeed
    from e3b200 import optim

    class F:
        data, n_graphs, seed = "synthetic", 48, 0
    config = configs.config_energy()
    loss_coeffs = dict(config.loss_coeffs.items()) if hasattr(config.loss_coeffs, "items") else dict(config.loss_coeffs)
    keys = list(loss_coeffs)
    data = train.load_data(F, config, keys)
    r_max = float(config.model_config.r_max)
    dev = torch.device("cuda")

    def run(feed):
        setSeed(0)
        model = build(config.model_config).to(dev).train()
        opt = optim.FlatAdam(model, lr=1e-3)
        losses = [float(train.train_step(model, opt, b, keys, loss_coeffs)[0]) for b in feed]
        return losses, opt.param.clone()

    pipe = train.make_pipeline(data, keys, 8, r_max, 0, 0, 1, dev)
    order = list(train.make_pipeline(data, keys, 8, r_max, 0, 0, 1, dev).batches_of_indices())[:3]
    it = pipe.endless()
    a_losses, a_params = run(next(it) for _ in range(3))

    attrs = {k: train.ATTRS[k] for k in data if k in train.ATTRS}
    ds = CondensedDataset(data=data, attrs=attrs)

    def direct():
        for g in order:
            b = _direct(ds, g, attrs).to(dev)
            d, a = computeEdgeIndex(b.data, b.attrs, r_max=r_max)
            b.attrs.update(a)
            b.update(d)
            yield Batch(b.attrs, **b.data)
    b_losses, b_params = run(direct())
    assert a_losses == b_losses and torch.equal(a_params, b_params)
    assert a_losses[-1] != a_losses[0]
