"""e3b200.optim.FlatAdam (flat parameter / gradient buffers, fused Adam + EMA kernel) against
torch.optim.Adam and a plain exponential moving average."""
import copy

import pytest
import torch

import harness
import product_harness
from e3b200 import optim, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _mlp(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(7, 33), torch.nn.SiLU(), torch.nn.Linear(33, 5)).to(DEV)


@pytest.mark.parametrize("clip,wd", [(None, 0.0), (0.05, 0.0), (None, 0.01)])
def test_flat_adam_matches_torch_adam_and_ema(clip, wd):
    ref, ours = _mlp(0), _mlp(0)
    opt_ref = torch.optim.Adam(ref.parameters(), lr=3e-3, weight_decay=wd)
    ema_ref = [p.detach().clone() for p in ref.parameters()]
    opt = optim.FlatAdam(ours, lr=3e-3, weight_decay=wd, ema_decay=0.99)
    g = torch.Generator().manual_seed(1)
    for step in range(1, 7):
        x = torch.randn(64, 7, generator=g).to(DEV)
        y = torch.randn(64, 5, generator=g).to(DEV)
        for model, zero in ((ref, lambda: opt_ref.zero_grad()), (ours, opt.zero_grad)):
            zero()
            ((model(x) - y) ** 2).mean().backward()
        if clip is not None:
            torch.nn.utils.clip_grad_norm_(ref.parameters(), clip)
        opt_ref.step()
        decay = min(0.99, (1 + step) / (10 + step))
        for e, p in zip(ema_ref, ref.parameters()):
            e.mul_(decay).add_(p.detach(), alpha=1 - decay)
        opt.step(max_grad_norm=clip)
        for a, b in zip(ours.parameters(), ref.parameters()):
            assert harness.rel_err(a, b) < 2e-6, step
    flat_ema = torch.cat([e.reshape(-1) for e in ema_ref])
    assert harness.rel_err(opt.ema, flat_ema) < 2e-6
    # parameters are views of the flat buffer; state_dict round-trips; EMA swap in / out
    assert all(p.data_ptr() >= opt.param.data_ptr() for p in ours.parameters())
    before = copy.deepcopy(ours.state_dict())
    opt.ema_swap_in()
    assert harness.rel_err(torch.cat([p.reshape(-1) for p in ours.parameters()]), flat_ema) < 2e-6
    opt.ema_swap_out()
    for k, v in ours.state_dict().items():
        assert torch.equal(v, before[k])


def test_flat_adam_skips_nonfinite_step():
    """a skipped step leaves parameters and moments alone, does NOT advance Adam's step count (torch.optim.Adam is
    simply not called by the reference loop, run/sde_utils.py:240-246) and still moves the moving average"""
    ref, m = _mlp(3), _mlp(3)
    opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-2)
    opt = optim.FlatAdam(m, lr=1e-2, ema_decay=0.9, ema_use_num_updates=False)
    x = torch.randn(8, 7, device=DEV)
    opt.zero_grad()
    m(x).sum().backward()
    opt.step()                                               # update 1
    opt_ref.zero_grad()
    ref(x).sum().backward()
    opt_ref.step()
    opt.zero_grad()
    m(x).sum().backward()
    opt.grad[3] = float("nan")
    before, ema_before, avg_before = opt.param.clone(), opt.ema.clone(), opt.exp_avg.clone()
    opt.step(skip_nonfinite=True)                            # skipped
    assert torch.equal(opt.param, before) and torch.equal(opt.exp_avg, avg_before)
    assert opt.applied_steps == 1 and opt.n_steps == 2
    assert harness.rel_err(opt.ema, 0.9 * ema_before + 0.1 * before) < 1e-6 and not torch.equal(opt.ema, ema_before)
    opt.zero_grad()
    m(x).sum().backward()
    opt.step(skip_nonfinite=True)                            # update 2: bias corrections of step 2, not 3
    opt_ref.zero_grad()
    ref(x).sum().backward()
    opt_ref.step()
    assert opt.applied_steps == 2
    for a, b in zip(m.parameters(), ref.parameters()):
        assert harness.rel_err(a, b) < 2e-6


def test_flat_adam_state_round_trip():
    a, b = _mlp(5), _mlp(6)
    oa, ob = optim.FlatAdam(a, lr=1e-2, ema_decay=0.99), optim.FlatAdam(b, lr=1e-3, ema_decay=0.99)
    x = torch.randn(8, 7, device=DEV)
    for _ in range(3):
        oa.zero_grad()
        a(x).sum().backward()
        oa.step()
    ob.load_state_dict(oa.state_dict())
    assert ob.n_steps == 3 and ob.applied_steps == 3 and ob.lr == 1e-2
    for o, model in ((oa, a), (ob, b)):
        o.zero_grad()
        model(x).sum().backward()
        o.step()
    assert torch.equal(oa.param, ob.param) and torch.equal(oa.ema, ob.ema)
    sd = oa.ema_state_dict(a)
    flat = torch.cat([sd[k].reshape(-1) for k, _ in a.named_parameters()])
    assert torch.equal(flat, oa.ema)


def test_flat_adam_drives_the_fused_blocks():
    """energy training through the fused interaction blocks: the packed tensor-core weights follow the flat updates"""
    meta = {"config": "config_energy", "seed": 5}
    inputs = synthetic.qm9_like(8, seed=4, n_min=3, n_max=9)
    model = product_harness.build_product(meta, torch.float32, DEV).train()
    with torch.no_grad():
        target = product_harness.run_product(model, inputs, torch.float32, DEV, pre_edge={"r_max": 4.0})["total_energy"] + 1.0
    opt = optim.FlatAdam(model, lr=1e-3, ema_decay=0.99)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        out = product_harness.run_product(model, inputs, torch.float32, DEV, pre_edge={"r_max": 4.0})
        loss = ((out["total_energy"] - target) ** 2).mean()
        loss.backward()
        opt.all_reduce()
        opt.step(max_grad_norm=10.0)
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0], losses
