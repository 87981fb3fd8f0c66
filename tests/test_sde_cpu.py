"""VP-SDE / PC sampler host logic (e3_layers.run) against the oracle restatement (oracle/ref_sde.py) with
identical recorded noise, on the CPU: the model underneath runs on the TEST-ONLY kernel stand-ins.
Closed-form checks pin the oracle's SDE pieces themselves (the reference ships no tests)."""
import math

import pytest
import torch

import product_harness
import torch_emulation
from e3_layers.run import VPSDE, LangevinCorrector, EulerMaruyamaPredictor, get_pc_sampler, get_sde_loss_fn
from e3b200 import synthetic
from sde_harness import Noise, oracle_data, oracle_model_fn, product_batch, ref_sde


def test_vpsde_closed_forms():
    sde = ref_sde.VPSDE({"pos": 3})
    data = {"t": torch.tensor([[0.3], [1.0]], dtype=torch.float64), "_node_segment": torch.tensor([0, 0, 1])}
    std = sde.std(data)
    for row, t in zip((0, 2), (0.3, 1.0)):
        integral = 0.1 * t + 0.5 * (20 - 0.1) * t * t            # int_0^t beta(s) ds
        assert abs(float(std[row]) - math.sqrt(1 - math.exp(-integral))) < 1e-12
    assert abs(float(sde.alphas[0]) - (1 - 0.1 / 1000)) < 1e-7 and abs(float(sde.alphas[-1]) - (1 - 20 / 1000)) < 1e-7
    # one forward Euler-Maruyama step with zero noise is the pure drift
    x0 = torch.ones(3, 3, dtype=torch.float64)
    d = {"pos": x0.clone(), **data}
    sde.sde_step(d, 1e-3, iter([torch.zeros(3, 3, dtype=torch.float64)]))
    beta = 0.1 + torch.tensor([0.3, 0.3, 1.0], dtype=torch.float64).view(-1, 1) * 19.9
    assert torch.allclose(d["pos"], x0 - 0.5 * beta * x0 * 1e-3, atol=1e-15)


def test_pc_sampler_and_loss_match_oracle_fp64(monkeypatch):
    torch_emulation.patch(monkeypatch)
    meta = {"config": "config_diffusion", "seed": 2}
    inputs = synthetic.diffusion_like(3, seed=1, n_min=3, n_max=6)
    N = inputs["pos"].shape[0]
    iters = 2
    # --- oracle
    osde = ref_sde.VPSDE({"pos": 3}, N=50)
    ref, nfe = ref_sde.pc_sampler(osde, oracle_model_fn(meta, inputs), oracle_data(inputs), snr=0.16, n_steps=1,
                                  noise=Noise((N, 3), 1 + 2 * iters, seed=7), max_iterations=iters)
    # --- product (CPU tensors, emulated kernels, eager)
    model = product_harness.build_product(meta, torch.float64, "cpu")
    sde = VPSDE({"pos": 3}, N=50)
    sde.randn_like = Noise((N, 3), 1 + 2 * iters, seed=7).randn_like
    sampler = get_pc_sampler(sde, EulerMaruyamaPredictor, LangevinCorrector, lambda b: b, snr=0.16, n_steps=1,
                             max_iterations=iters, graph=False)
    torch.set_default_dtype(torch.float64)
    try:
        out, nfe2 = sampler(model, product_batch(inputs, torch.float64, "cpu"))
    finally:
        torch.set_default_dtype(torch.float32)
    assert nfe == nfe2 == 2 * iters
    assert float((out["pos"] - ref["pos"]).abs().max()) < 1e-9 * float(ref["pos"].abs().max())
    # --- score-matching loss with recorded t and z
    t = torch.tensor([0.2, 0.5, 0.9], dtype=torch.float64)
    ref_loss = ref_sde.sde_loss(osde, oracle_model_fn(meta, inputs), oracle_data(inputs), t, Noise((N, 3), 1, seed=9))
    sde.randn_like = Noise((N, 3), 1, seed=9).randn_like
    monkeypatch.setattr(torch, "rand", lambda *a, **k: ((t - 1e-5) / (1 - 1e-5)).to(k.get("device", "cpu")))
    torch.set_default_dtype(torch.float64)
    try:
        loss, _ = get_sde_loss_fn(sde, train=False)(model, product_batch(inputs, torch.float64, "cpu"))
    finally:
        torch.set_default_dtype(torch.float32)
    assert abs(float(loss) - float(ref_loss)) < 1e-9 * abs(float(ref_loss))


def test_step_fn_follows_the_reference_quirks(monkeypatch):
    """get_step_fn (reference sde_utils.py:204-257): no optimiser step at step 0, gradient accumulation over
    grad_acc steps, EMA updated every call, evaluation runs on the EMA weights and restores the live ones,
    the grad_sync hook (flat gradient all-reduce) runs after every backward."""
    from e3_layers.run import ExponentialMovingAverage, get_step_fn

    torch_emulation.patch(monkeypatch)
    meta = {"config": "config_diffusion", "seed": 2}
    inputs = synthetic.diffusion_like(2, seed=1, n_min=3, n_max=5)
    model = product_harness.build_product(meta, torch.float64, "cpu").train()
    sde = VPSDE({"pos": 3}, N=50)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    state = {"model": model, "optimizer": opt, "ema": ExponentialMovingAverage(model.parameters(), decay=0.5), "step": 0}
    synced = []
    step = get_step_fn(sde, train=True, optimizer=opt, grad_clid_norm=1.0, grad_acc=2, grad_sync=lambda: synced.append(1))
    flat = lambda: torch.cat([p.detach().reshape(-1).clone() for p in model.parameters()])
    torch.set_default_dtype(torch.float64)
    try:
        p0 = flat()
        batch = lambda: product_batch({k: v for k, v in inputs.items() if k != "t"}, torch.float64, "cpu")
        loss0, parts = step(state, batch())                       # step 0: backward only
        assert torch.equal(flat(), p0) and state["step"] == 1 and loss0 == loss0 and "pos" in parts and "total" in parts
        step(state, batch())                                      # step 1: 1 % 2 != 0 -> still accumulating
        assert torch.equal(flat(), p0)
        step(state, batch())                                      # step 2: optimiser step with the accumulated gradient
        p2 = flat()
        assert not torch.equal(p2, p0) and len(synced) == 3
        ema = torch.cat([s.reshape(-1) for s in state["ema"].shadow])
        assert not torch.equal(ema, p2) and not torch.equal(ema, p0)      # lags behind the live weights
        val, _ = get_step_fn(sde, train=False)(state, batch())
        assert val == val and torch.equal(flat(), p2)             # evaluation on the EMA weights leaves the live ones alone
    finally:
        torch.set_default_dtype(torch.float32)


def test_ema_decay_warms_up_like_the_reference():
    """score_sde_pytorch models/ema.py (what reference train.py:103 builds): use_num_updates=True by default, decay =
    min(decay, (1 + n) / (10 + n)); parameters without requires_grad are not tracked; state round-trips."""
    from e3_layers.run import ExponentialMovingAverage

    w = torch.nn.Parameter(torch.zeros(3))
    frozen = torch.nn.Parameter(torch.ones(2), requires_grad=False)
    ema = ExponentialMovingAverage([w, frozen], decay=0.9999)
    assert len(ema.shadow) == 1
    ref = torch.zeros(3)
    for n in range(1, 30):
        with torch.no_grad():
            w += 1.0
        ema.update([w, frozen])
        d = min(0.9999, (1 + n) / (10 + n))
        ref = d * ref + (1 - d) * w.detach()
        assert abs(ema.current_decay() - d) < 1e-15
        assert torch.allclose(ema.shadow[0], ref, atol=1e-6)
    assert float(ema.shadow[0][0]) > 10.0            # a fixed 0.9999 would still sit at ~0.04
    fixed = ExponentialMovingAverage([w], decay=0.9999, use_num_updates=False)
    fixed.update([w])
    assert fixed.current_decay() == 0.9999
    other = ExponentialMovingAverage([torch.nn.Parameter(torch.zeros(3)), frozen], decay=0.5)
    other.load_state_dict(ema.state_dict())
    assert other.num_updates == 29 and other.decay == 0.9999 and torch.equal(other.shadow[0], ema.shadow[0])
    ema.store([w, frozen])
    ema.copy_to([w, frozen])
    assert torch.equal(w.detach(), ema.shadow[0]) and torch.equal(frozen, torch.ones(2))
    ema.restore([w, frozen])
    assert float(w[0]) == 29.0
