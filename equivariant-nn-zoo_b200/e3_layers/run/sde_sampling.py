"""Predictor-corrector sampling (reference ``e3_layers/run/sde_sampling.py:14-246``): Euler-Maruyama
predictor on the reverse SDE, Langevin corrector, ``get_pc_sampler``.

B200 specifics: one sampler iteration (corrector steps + predictor step, i.e. ``n_steps + 1`` score
evaluations and their position updates, noise drawn on the device) is captured ONCE as a CUDA graph and
replayed ``sde.N`` times when the model does not rebuild its neighbour list (``config_diffusion``:
complete graphs from the dataset, fixed topology); models whose first layer recomputes the edges
(``config_diffusion_CA``) run the same code eagerly."""
import abc

import torch

from .sde_utils import VPSDE, get_score_fn

_CORRECTORS, _PREDICTORS = {}, {}


def _register(table, name):
    def deco(cls):
        key = name or cls.__name__
        if key in table:
            raise ValueError(f"Already registered model with name: {key}")
        table[key] = cls
        return cls
    return deco


def register_predictor(cls=None, *, name=None):
    return _register(_PREDICTORS, name)(cls) if cls is not None else _register(_PREDICTORS, name)


def register_corrector(cls=None, *, name=None):
    return _register(_CORRECTORS, name)(cls) if cls is not None else _register(_CORRECTORS, name)


def get_predictor(name):
    return _PREDICTORS[name]


def get_corrector(name):
    return _CORRECTORS[name]


class Predictor(abc.ABC):
    def __init__(self, sde, score_fn):
        self.sde, self.score_fn = sde, score_fn
        self.rsde = sde.reverse(score_fn)

    @abc.abstractmethod
    def update_fn(self, batch):
        ...


class Corrector(abc.ABC):
    def __init__(self, sde, score_fn, snr, n_steps):
        self.sde, self.score_fn, self.snr, self.n_steps = sde, score_fn, snr, n_steps

    @abc.abstractmethod
    def update_fn(self, batch):
        ...


@register_predictor(name="euler_maruyama")
class EulerMaruyamaPredictor(Predictor):
    def update_fn(self, batch):
        return self.rsde.sde(batch)


@register_predictor(name="none")
class NonePredictor(Predictor):
    def __init__(self, sde, score_fn):
        pass

    def update_fn(self, batch):
        return batch


def _drop_stale_geometry(batch, rebuild_edges):
    """positions moved: edge vectors are stale; the edge list too when the model recomputes it"""
    for key in ("edge_vector", "edge_length") + (("edge_index", "_n_edges", "_edge_segment") if rebuild_edges else ()):
        batch.pop(key)


@register_corrector(name="langevin")
class LangevinCorrector(Corrector):
    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__(sde, score_fn, snr, n_steps)
        if not isinstance(sde, VPSDE):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")

    def update_fn(self, batch):
        """the reference (sde_sampling.py:121-140) evaluates the score n_steps times on the UNCHANGED batch (it writes
        the positions back only after the loop), so every inner step sees the same score: it is evaluated once here."""
        sde = self.sde
        key = next(iter(sde.irreps))
        x = batch[key]
        t = batch["t"].reshape(-1, 1)[batch.nodeSegment()]
        timestep = (t * (sde.N - 1) / sde.T).long()
        alpha = sde.alphas_on(t.device)[timestep].to(x.dtype)
        result = self.score_fn(batch)
        grad = result[sde.score_key(result, key)]
        for _ in range(self.n_steps):
            noise = sde.randn_like(x)
            grad_norm = torch.norm(grad.reshape(grad.shape[0], -1), dim=-1).mean()
            noise_norm = torch.norm(noise.reshape(noise.shape[0], -1), dim=-1).mean()
            step_size = (self.snr * noise_norm / grad_norm) ** 2 * 2 * alpha
            x_mean = x + step_size * grad
            x = x_mean + torch.sqrt(step_size * 2) * noise
        batch[key] = x
        return batch


@register_corrector(name="none")
class NoneCorrector(Corrector):
    def __init__(self, sde, score_fn, snr, n_steps):
        pass

    def update_fn(self, batch):
        return batch


def shared_predictor_update_fn(batch, sde, model, predictor, continuous):
    score_fn = get_score_fn(sde, model, train=False)
    obj = NonePredictor(sde, score_fn) if predictor is None else predictor(sde, score_fn)
    return obj.update_fn(batch)


def shared_corrector_update_fn(batch, sde, model, corrector, continuous, snr, n_steps):
    score_fn = get_score_fn(sde, model, train=False)
    obj = NoneCorrector(sde, score_fn, snr, n_steps) if corrector is None else corrector(sde, score_fn, snr, n_steps)
    return obj.update_fn(batch)


def _model_rebuilds_edges(model):
    layers = getattr(model, "layers", None)
    if layers is None and hasattr(model, "func"):
        layers = getattr(model.func, "layers", None)
    for key, layer in (layers or []):
        fn = getattr(layer, "func", layer)
        if getattr(fn, "__name__", "") == "computeEdgeIndex":
            return True
    return False


def get_pc_sampler(sde, predictor, corrector, inverse_scaler, snr, n_steps=1, continuous=False, eps=1e-3,
                   graph=True, max_iterations=None):
    """-> pc_sampler(model, batch) -> (samples, number of score evaluations).  ``graph``: replay one captured
    CUDA graph per iteration when the topology is fixed; ``max_iterations`` truncates the loop (tests / benchmarks)."""

    def iteration(model, batch, rebuild):
        batch = shared_corrector_update_fn(batch, sde=sde, model=model, corrector=corrector, continuous=continuous, snr=snr,
                                           n_steps=n_steps)
        _drop_stale_geometry(batch, rebuild)
        batch = shared_predictor_update_fn(batch, sde=sde, model=model, predictor=predictor, continuous=continuous)
        _drop_stale_geometry(batch, rebuild)
        return batch

    def pc_sampler(model, batch):
        batch = batch.clone()
        batch.attrs["t"] = ("graph", "1x0e")
        batch = sde.prior_sampling(batch)
        dev = batch["_n_nodes"].device
        keys = list(sde.irreps)
        dtype = batch[keys[0]].dtype
        timesteps = torch.linspace(sde.T, eps, sde.N, device=dev, dtype=dtype)
        n_iter = sde.N if max_iterations is None else min(sde.N, max_iterations)
        rebuild = _model_rebuilds_edges(model)
        use_graph = bool(graph and dev.type == "cuda" and not rebuild and "edge_index" in batch)
        with torch.no_grad():
            # the sampler owns 't' (sde_sampling.py:231-236 sets it every iteration): dataset batches do not carry it,
            # so it is assigned BEFORE the set of graph inputs is recorded
            batch["t"] = timesteps[0].expand(len(batch)).reshape(-1, 1).clone()
            inputs = set(batch.keys())
            if use_graph:
                static_t = batch["t"]
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                saved = {k: batch[k].clone() for k in keys}
                with torch.cuda.stream(side):                 # warm-up off the capture (lazy plans, weight packs)
                    iteration(model, batch, rebuild)
                torch.cuda.current_stream().wait_stream(side)
                static_x = {k: saved[k] for k in keys}
                for k in list(batch.keys()):                  # drop everything the warm-up iteration left behind
                    if k not in inputs:
                        batch.pop(k)
                for k in keys:
                    batch[k] = static_x[k]
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    out = iteration(model, batch, rebuild)
                    for k in keys:
                        static_x[k].copy_(out[k])             # next replay starts from this iteration's result
                for i in range(n_iter):
                    static_t.copy_(timesteps[i].expand_as(static_t))
                    g.replay()
                for k in keys:
                    batch[k] = static_x[k].clone()
            else:
                for i in range(n_iter):
                    batch["t"] = timesteps[i].expand(len(batch)).reshape(-1, 1).clone()
                    batch = iteration(model, batch, rebuild)
        return inverse_scaler(batch), n_iter * (n_steps + 1)

    return pc_sampler
