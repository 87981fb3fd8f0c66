"""TEST INFRASTRUCTURE -- CPU restatement of the reference's VP-SDE, score function and
predictor-corrector sampler (``e3_layers/run/sde_utils.py:32-125,190-201`` and
``e3_layers/run/sde_sampling.py:97-246``) on plain dicts of torch tensors.  Only tests, smoke() and
the cpu_baseline leg of bench.py may import this package.

``model_fn(data) -> result`` evaluates a score model on ``data`` (keys: the diffused key, ``t`` [G,1],
``_node_segment`` [N]).  ``noise`` is an iterator of pre-drawn standard normal tensors, consumed in the
order the reference calls ``torch.randn_like`` (corrector noise, then predictor noise, per iteration), so
that another implementation can be fed the identical noise.

Where the reference is inconsistent at HEAD this file does what its code would do
with the inconsistency removed in the most literal way: ``t`` is a per-graph column [G,1] (the sampler
sets a 0-dim ``t`` that ``marginal`` then indexes per node, sde_sampling.py:233 vs sde_utils.py:55); the
score key is ``score_<key>`` or ``score`` (sde_utils.py:196 vs config_diffusion.py:112)."""
import math

import torch


class VPSDE:
    """sde_utils.py:32-125"""

    def __init__(self, diffusion_keys, beta_min=0.1, beta_max=20, N=1000):
        self.beta_0, self.beta_1, self.N = beta_min, beta_max, N
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)         # :43
        self.alphas = 1.0 - self.discrete_betas                                       # :44
        self.irreps = diffusion_keys
        self.T = 1

    def std(self, data):
        """marginal(..., return_std=True), :54-59"""
        t = data["t"].reshape(-1, 1)[data["_node_segment"]]
        lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return torch.sqrt(1.0 - torch.exp(2.0 * lmc))

    def marginal(self, data, noise):
        """:54-67 -- perturbs the diffused keys in place, returns (z, std)"""
        t = data["t"].reshape(-1, 1)[data["_node_segment"]]
        lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        std = torch.sqrt(1.0 - torch.exp(2.0 * lmc))
        zs = {}
        for key in self.irreps:
            z = next(noise).to(data[key])
            data[key] = torch.exp(lmc) * data[key] + std * z
            zs[key] = z
        return zs, std

    def sde_step(self, data, dt, noise):
        """:69-82"""
        t = data["t"].reshape(-1, 1)[data["_node_segment"]]
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        diffusion = torch.sqrt(beta_t)
        for key in self.irreps:
            x = data[key]
            x_mean = x + (-0.5 * beta_t * x) * dt
            data[key] = x_mean + diffusion * math.sqrt(abs(dt)) * next(noise).to(x)
        return data


def _score_key(result, key):
    return f"score_{key}" if f"score_{key}" in result else "score"


def score_fn(sde, model_fn, data):
    """get_score_fn, sde_utils.py:190-201: score = -model_output / std - x"""
    result = dict(model_fn(data))
    std = sde.std(data)
    for key in sde.irreps:
        name = _score_key(result, key)
        result[name] = -result[name] / std - data[key]
    return result


def reverse_step(sde, model_fn, data, noise):
    """RSDE.sde, sde_utils.py:108-121 (the Euler-Maruyama predictor, sde_sampling.py:97-104)"""
    scores = score_fn(sde, model_fn, data)
    t = data["t"].reshape(-1, 1)[data["_node_segment"]]
    beta_t = sde.beta_0 + t * (sde.beta_1 - sde.beta_0)
    dt = -1.0 / sde.N
    sde.sde_step(data, dt, noise)
    for key in sde.irreps:
        data[key] = data[key] - dt * beta_t * scores[_score_key(scores, key)]
    return data


def langevin_step(sde, model_fn, data, snr, n_steps, noise):
    """LangevinCorrector.update_fn, sde_sampling.py:121-140 (score evaluated on the batch as it entered)"""
    key = next(iter(sde.irreps))
    x = data[key]
    t = data["t"].reshape(-1, 1)[data["_node_segment"]]
    timestep = (t * (sde.N - 1) / sde.T).long()
    alpha = sde.alphas[timestep].to(x.dtype)
    for _ in range(n_steps):
        result = score_fn(sde, model_fn, data)
        grad = result[_score_key(result, key)]
        z = next(noise).to(x)
        grad_norm = torch.norm(grad.reshape(grad.shape[0], -1), dim=-1).mean()
        noise_norm = torch.norm(z.reshape(z.shape[0], -1), dim=-1).mean()
        step_size = (snr * noise_norm / grad_norm) ** 2 * 2 * alpha
        x_mean = x + step_size * grad
        x = x_mean + torch.sqrt(step_size * 2) * z
    data[key] = x
    return data


def pc_sampler(sde, model_fn, data, snr, n_steps, noise, eps=1e-3, max_iterations=None):
    """get_pc_sampler.pc_sampler, sde_sampling.py:215-246: prior sample, then corrector + predictor per time step"""
    data = dict(data)
    for key, dim in sde.irreps.items():
        data[key] = next(noise).to(data[key])                       # prior_sampling, sde_utils.py:84-87
    timesteps = torch.linspace(sde.T, eps, sde.N, dtype=data[next(iter(sde.irreps))].dtype)
    n_iter = sde.N if max_iterations is None else min(sde.N, max_iterations)
    G = data["_n_nodes"].shape[0]
    for i in range(n_iter):
        data["t"] = timesteps[i].expand(G).reshape(-1, 1).clone()
        data = langevin_step(sde, model_fn, data, snr, n_steps, noise)
        data = reverse_step(sde, model_fn, data, noise)
    return data, n_iter * (n_steps + 1)


def sde_loss(sde, model_fn, data, t, noise, reduce_mean=True):
    """get_sde_loss_fn.loss_fn, sde_utils.py:147-172 with the random draws (t, z) passed in"""
    data = dict(data)
    data["t"] = t.reshape(-1, 1)
    zs, std = sde.marginal(data, noise)
    scores = score_fn(sde, model_fn, data)
    total = 0.0
    for key in sde.irreps:
        loss = torch.square(scores[_score_key(scores, key)] * std + zs[key])
        loss = loss.reshape(loss.shape[0], -1)
        loss = loss.mean(dim=-1) if reduce_mean else 0.5 * loss.sum(dim=-1)
        total = total + loss.mean()
    return total
