"""Variance-preserving SDE, score function and score-matching loss (reference
``e3_layers/run/sde_utils.py:32-201``), operating on ``Batch`` objects.

Differences from the reference, all forced by its state at HEAD: the model's score
output is looked up as ``score_<key>`` and, failing that, ``score`` (``config_diffusion`` names its head
``score`` while ``sde_utils.py:196`` reads ``score_pos``); noise comes from ``sde.randn_like`` so that tests can
inject the same noise into the oracle; tensors are created on the batch's device (``prior_sampling`` in the
reference draws on the CPU and moves)."""
import math

import torch


class VPSDE:
    def __init__(self, diffusion_keys, beta_min=0.1, beta_max=20, N=1000):
        self.beta_0, self.beta_1, self.N = beta_min, beta_max, N
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1.0 - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_1m_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)
        self.irreps = diffusion_keys              # {diffused key: dim}
        self.randn_like = torch.randn_like        # hook: tests replace it to inject recorded noise
        self._alphas_dev = {}

    def alphas_on(self, device):
        """discrete alphas on `device`, copied once (a host-to-device copy cannot sit inside a captured iteration)"""
        a = self._alphas_dev.get(device)
        if a is None:
            a = self._alphas_dev[device] = self.alphas.to(device)
        return a

    @property
    def T(self):
        return 1

    def score_key(self, result, key):
        name = f"score_{key}"
        return name if name in result else "score"

    def marginal(self, batch, return_std=False):
        t = batch["t"].reshape(-1, 1)[batch.nodeSegment()]
        log_mean_coeff = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        std = torch.sqrt(1.0 - torch.exp(2.0 * log_mean_coeff))
        if return_std:
            return std
        zs = {}
        for key in self.irreps:
            mean = torch.exp(log_mean_coeff) * batch[key]
            z = self.randn_like(batch[key])
            batch[key] = mean + std * z
            zs[key] = z
        return batch, {"zs": zs, "std": std}

    def sde(self, batch, dt=None):
        """one Euler-Maruyama step of the forward SDE with step dt (negative dt: backwards in time)"""
        if dt is None:
            dt = 1.0 / self.N
        t = batch["t"].reshape(-1, 1)[batch.nodeSegment()]       # per node (the reference broadcasts a scalar t)
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        diffusion = torch.sqrt(beta_t)
        for key in self.irreps:
            x = batch[key]
            x_mean = x + (-0.5 * beta_t * x) * dt
            batch[key] = x_mean + diffusion * math.sqrt(abs(dt)) * self.randn_like(x)
        return batch

    def prior_sampling(self, batch):
        n = batch["_n_nodes"]
        ref = next((batch[k] for k in self.irreps if k in batch), None)
        for key, dim in self.irreps.items():
            like = torch.zeros(ref.shape[0] if ref is not None else int(n.sum()), dim, device=n.device,
                               dtype=ref.dtype if ref is not None else torch.get_default_dtype())
            batch[key] = self.randn_like(like)
        return batch

    def reverse(self, score_fn):
        """the reverse-time SDE: one step = forward-SDE step with dt = -1/N plus the score drift"""
        outer = self

        class RSDE:
            N, T = outer.N, outer.T

            def sde(self, batch):
                scores = score_fn(batch)
                t = batch["t"].reshape(-1, 1)[batch.nodeSegment()]
                beta_t = outer.beta_0 + t * (outer.beta_1 - outer.beta_0)
                dt = -1.0 / outer.N
                picked = {key: scores[outer.score_key(scores, key)] for key in outer.irreps}
                batch = outer.sde(batch, dt)
                for key in outer.irreps:
                    batch[key] = batch[key] - dt * beta_t * picked[key]           # diffusion^2 = beta_t
                return batch

        return RSDE()


def get_score_fn(sde, model, train=False):
    def score_fn(batch):
        model.train() if train else model.eval()
        result = model(batch)
        std = sde.marginal(batch, return_std=True)
        for key in sde.irreps:
            name = sde.score_key(result, key)
            result[name] = -result[name] / std - batch[key]
        return result

    return score_fn


def get_sde_loss_fn(sde, train, reduce_mean=True, continuous=True, likelihood_weighting=True, eps=1e-5):
    reduce_op = torch.mean if reduce_mean else (lambda *a, **k: 0.5 * torch.sum(*a, **k))

    def loss_fn(model, batch):
        t = torch.rand(len(batch), device=batch["_n_nodes"].device) * (sde.T - eps) + eps
        score_fn = get_score_fn(sde, model, train)
        perturbed = batch.clone()
        perturbed.attrs["t"] = ("graph", "1x0e")
        perturbed["t"] = t
        perturbed, misc = sde.marginal(perturbed)
        scores = score_fn(perturbed)
        losses = {}
        for key in sde.irreps:
            loss = torch.square(scores[sde.score_key(scores, key)] * misc["std"] + misc["zs"][key])
            losses[key] = torch.mean(reduce_op(loss.reshape(loss.shape[0], -1), dim=-1))
        total = sum(losses.values())
        losses["total"] = total
        return total, losses

    return loss_fn


class ExponentialMovingAverage:
    """The moving average the reference's diffusion loop builds (``train.py:103`` ->
    score_sde_pytorch ``models/ema.py``): only parameters with ``requires_grad`` are tracked, and with
    ``use_num_updates`` (the default there) the decay warms up as ``min(decay, (1 + n) / (10 + n))`` so that the
    average leaves the random initial weights within a few hundred steps even at ``ema_rate = 0.9999``."""

    def __init__(self, parameters, decay=0.999, use_num_updates=True):
        if not 0.0 <= decay <= 1.0:
            raise ValueError("decay must be between 0 and 1")
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        self.shadow = [p.detach().clone() for p in parameters if p.requires_grad]
        self.backup = None

    def current_decay(self):
        if self.num_updates is None:
            return self.decay
        return min(self.decay, (1.0 + self.num_updates) / (10.0 + self.num_updates))

    @torch.no_grad()
    def update(self, parameters):
        if self.num_updates is not None:
            self.num_updates += 1
        ps = [p.detach() for p in parameters if p.requires_grad]
        torch._foreach_lerp_(self.shadow, ps, 1.0 - self.current_decay())       # one multi-tensor kernel

    def store(self, parameters):
        self.backup = [p.detach().clone() for p in parameters if p.requires_grad]

    @torch.no_grad()
    def copy_to(self, parameters):
        for s, p in zip(self.shadow, [p for p in parameters if p.requires_grad]):
            p.copy_(s)

    @torch.no_grad()
    def restore(self, parameters):
        for b, p in zip(self.backup, [p for p in parameters if p.requires_grad]):
            p.copy_(b)

    def state_dict(self):
        return {"decay": self.decay, "num_updates": self.num_updates, "shadow_params": [s.clone() for s in self.shadow]}

    def load_state_dict(self, state):
        self.decay, self.num_updates = state["decay"], state["num_updates"]
        for s, v in zip(self.shadow, state["shadow_params"]):
            s.copy_(v.to(s.device))


def get_step_fn(sde, train, optimizer=None, reduce_mean=False, continuous=True, likelihood_weighting=False,
                grad_clid_norm=None, grad_acc=1, grad_sync=None):
    """one training / evaluation step over ``state = {model, optimizer, ema, step}`` (sde_utils.py:204-257).
    ``grad_sync``: called after the backward pass (the flat gradient all-reduce that stands in for the reference's DDP)"""
    loss_fn = get_sde_loss_fn(sde, train, reduce_mean=reduce_mean, continuous=True, likelihood_weighting=likelihood_weighting)

    def step_fn(state, batch):
        model = state["model"]
        if train:
            opt = state["optimizer"]
            loss, losses = loss_fn(model, batch)
            loss.backward()
            if grad_sync is not None:
                grad_sync()
            if grad_clid_norm is not None:
                torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=grad_clid_norm)
            if state["step"] != 0 and state["step"] % grad_acc == 0:
                bad = torch.stack([(~torch.isfinite(p.grad)).any() for p in model.parameters() if p.grad is not None]).any()
                if not bool(bad):
                    opt.step()
                opt.zero_grad(set_to_none=True)
            state["step"] += 1
            state["ema"].update(model.parameters())
        else:
            ema = state["ema"]
            ema.store(model.parameters())
            ema.copy_to(model.parameters())
            with torch.no_grad():
                loss, losses = loss_fn(model, batch)
            ema.restore(model.parameters())
        return loss.item(), {k: v.item() for k, v in losses.items()}

    return step_fn
