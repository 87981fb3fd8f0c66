"""PC sampler and score-matching step on the GPU (e3_layers.run): product fp32 against the
fp64 oracle restatement with identical recorded noise; the CUDA-graph replay against the eager loop; the
protein model, whose first layer rebuilds the neighbour list every evaluation."""
import pytest
import torch

import harness
import product_harness
from e3_layers import configs
from e3_layers.data import Batch
from e3_layers.run import (VPSDE, EulerMaruyamaPredictor, ExponentialMovingAverage, LangevinCorrector, get_pc_sampler,
                           get_step_fn)
from e3_layers.utils import build
from e3b200 import synthetic
from param_init import reseed_parameters
from sde_harness import Noise, oracle_data, oracle_model_fn, product_batch, ref_sde

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_pc_sampler_against_oracle_recorded_noise():
    meta = {"config": "config_diffusion", "seed": 2}
    inputs = synthetic.diffusion_like(5, seed=1, n_min=3, n_max=9)
    N, iters = inputs["pos"].shape[0], 3
    osde = ref_sde.VPSDE({"pos": 3}, N=100)
    ref, _ = ref_sde.pc_sampler(osde, oracle_model_fn(meta, inputs), oracle_data(inputs), snr=0.16, n_steps=1,
                                noise=Noise((N, 3), 1 + 2 * iters, seed=7), max_iterations=iters)
    for dtype, tol in ((torch.float64, 1e-9), (torch.float32, 1e-4)):
        model = product_harness.build_product(meta, dtype, DEV)
        sde = VPSDE({"pos": 3}, N=100)
        sde.randn_like = Noise((N, 3), 1 + 2 * iters, seed=7).randn_like
        sampler = get_pc_sampler(sde, EulerMaruyamaPredictor, LangevinCorrector, lambda b: b, snr=0.16, n_steps=1,
                                 max_iterations=iters, graph=False)
        torch.set_default_dtype(dtype)
        try:
            out, nfe = sampler(model, product_batch(inputs, dtype, DEV))
        finally:
            torch.set_default_dtype(torch.float32)
        assert nfe == 2 * iters
        assert harness.rel_err(out["pos"], ref["pos"]) < tol, (dtype, harness.rel_err(out["pos"], ref["pos"]))


def test_graph_replay_equals_eager_loop():
    """one captured iteration replayed == the eager loop, with a deterministic stand-in for the noise"""
    meta = {"config": "config_diffusion", "seed": 4}
    inputs = synthetic.diffusion_like(16, seed=3)
    model = product_harness.build_product(meta, torch.float32, DEV)
    outs = []
    for graph in (False, True):
        sde = VPSDE({"pos": 3}, N=200)
        # deterministic stand-in for the noise: same values in both runs, capturable, distinct per atom
        sde.randn_like = lambda x: 1.2 * torch.sin(37.0 * x + 0.7 * torch.arange(x.shape[0], device=x.device,
                                                                                 dtype=x.dtype).view(-1, 1) + 1.3)
        sampler = get_pc_sampler(sde, EulerMaruyamaPredictor, LangevinCorrector, lambda b: b, snr=0.16, n_steps=1,
                                 max_iterations=6, graph=graph)
        out, nfe = sampler(model, product_batch(inputs, torch.float32, DEV))
        assert nfe == 12 and torch.isfinite(out["pos"]).all()
        outs.append(out["pos"])
    assert harness.rel_err(outs[1], outs[0]) < 1e-5
    # and with device-side random noise the replayed graph draws fresh numbers every iteration
    sde = VPSDE({"pos": 3}, N=200)
    sampler = get_pc_sampler(sde, EulerMaruyamaPredictor, LangevinCorrector, lambda b: b, snr=0.16, max_iterations=4)
    a, _ = sampler(model, product_batch(inputs, torch.float32, DEV))
    b, _ = sampler(model, product_batch(inputs, torch.float32, DEV))
    assert torch.isfinite(a["pos"]).all() and not torch.equal(a["pos"], b["pos"])


def test_graph_sampler_on_a_batch_without_t():
    """dataset batches (qm9 ...) carry no 't': the sampler assigns it itself (reference sde_sampling.py:231-236), also
    on the CUDA-graph path (the set of graph inputs was recorded before 't' existed -> KeyError)"""
    meta = {"config": "config_diffusion", "seed": 4}
    inputs = {k: v for k, v in synthetic.diffusion_like(8, seed=5).items() if k != "t"}
    model = product_harness.build_product(meta, torch.float32, DEV)
    outs = []
    for graph in (False, True):
        sde = VPSDE({"pos": 3}, N=100)
        sde.randn_like = lambda x: 0.9 * torch.cos(11.0 * x + 0.3 * torch.arange(x.shape[0], device=x.device,
                                                                                 dtype=x.dtype).view(-1, 1))
        sampler = get_pc_sampler(sde, EulerMaruyamaPredictor, LangevinCorrector, lambda b: b, snr=0.16, n_steps=1,
                                 max_iterations=4, graph=graph)
        batch = product_batch(inputs, torch.float32, DEV)
        assert "t" not in batch
        torch.manual_seed(0)
        out, nfe = sampler(model, batch)
        assert nfe == 8 and torch.isfinite(out["pos"]).all()
        outs.append(out["pos"])
    # the prior draw differs between the two runs unless seeded identically: both were seeded with 0 above
    assert harness.rel_err(outs[1], outs[0]) < 1e-5


def test_protein_sampler_rebuilds_edges_and_score_matching_step():
    cfg = configs.config_diffusion_CA()
    model = reseed_parameters(build(cfg.model_config), 1).to(DEV).eval()
    inputs = synthetic.protein_like(120, seed=2)
    inputs.pop("edge_index"), inputs.pop("_n_edges")
    sde = VPSDE(dict(cfg.diffusion_keys.items()) if hasattr(cfg.diffusion_keys, "items") else {"CA": 3}, N=100)
    sampler = get_pc_sampler(sde, EulerMaruyamaPredictor, LangevinCorrector, lambda b: b, snr=0.16, max_iterations=2)
    out, nfe = sampler(model, product_batch(inputs, torch.float32, DEV))
    assert nfe == 4 and out["CA"].shape == (120, 3) and torch.isfinite(out["CA"]).all()
    # score-matching training step (reference get_step_fn): loss is finite and parameters move
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    state = {"model": model, "optimizer": opt, "ema": ExponentialMovingAverage(model.parameters()), "step": 0}
    step = get_step_fn(sde, train=True, optimizer=opt, grad_clid_norm=1.0)
    before = [p.detach().clone() for p in model.parameters()]
    losses = [step(state, product_batch(inputs, torch.float32, DEV))[0] for _ in range(3)]
    assert all(l == l and l < float("inf") for l in losses)
    assert any(not torch.equal(a, b) for a, b in zip(before, model.parameters()))
    val, _ = get_step_fn(sde, train=False)(state, product_batch(inputs, torch.float32, DEV))
    assert val == val
