"""Helpers of ``e3_layers.utils`` needed by the model graph: the layer factory ``build``, kwarg
pruning, key remapping, path-existence test, the activation table and list editing helpers
(reference: ``e3_layers/utils/utils.py:9-156``), without e3nn / ml_collections."""
import inspect
import math

import numpy as np
import torch

from e3b200.irreps import Irrep, Irreps

from .config_dict import ConfigDict


def setSeed(seed):
    torch.manual_seed(seed)
    np.random.seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def tanhlu(x):
    return torch.tanh(x) * torch.abs(x)


def ShiftedSoftPlus(x):
    return torch.nn.functional.softplus(x) - math.log(2.0)


# name -> callable, same names as the reference's table (utils.py:78-84)
activations = {"abs": torch.abs, "tanh": torch.tanh, "ssp": ShiftedSoftPlus,
               "silu": torch.nn.functional.silu, "tanhlu": tanhlu}


def activation_name(fn_or_name):
    if isinstance(fn_or_name, str):
        return fn_or_name
    for name, fn in activations.items():
        if fn is fn_or_name:
            return name
    raise KeyError(f"unknown activation {fn_or_name}")


def tp_path_exists(irreps_in1, irreps_in2, ir_out):
    ir_out = Irrep(ir_out)
    a, b = Irreps(irreps_in1).simplify(), Irreps(irreps_in2).simplify()
    return any(ir_out in x.ir * y.ir for x in a for y in b)


def pruneArgs(_func=None, prefix="", **kwargs):
    """keep kwargs that start with `prefix_` (stripped) and that `_func` accepts"""
    if prefix:
        kwargs = {k[len(prefix) + 1:]: v for k, v in kwargs.items() if k.startswith(prefix)}
    if _func is None:
        return kwargs
    target = _func.__init__ if inspect.isclass(_func) else _func
    params = inspect.signature(target).parameters
    if any(p.kind is inspect.Parameter.VAR_KEYWORD for p in params.values()):
        return kwargs
    return {k: v for k, v in kwargs.items() if k in params}


def build(node, **kwargs):
    """Instantiate a layer from its config node: a mapping with a "module" entry (class or
    callable) plus constructor kwargs, or a bare callable."""
    if isinstance(node, (dict, ConfigDict)):
        merged = dict(kwargs)
        merged.update({k: node[k] for k in node.keys()})
        factory = merged.pop("module")
    elif isinstance(node, (list, tuple)):
        factory, merged = node[0], dict(kwargs)
    else:
        factory, merged = node, dict(kwargs)
    merged.pop("module", None)
    return factory(**pruneArgs(factory, **merged))


def keyMap(dic, key_mapping):
    """rename keys of a dict (or of a Data object's tensors and attrs); a target may be a list"""
    if not isinstance(dic, dict):
        return type(dic)(keyMap(dic.attrs, key_mapping), **keyMap(dic.data, key_mapping))
    out = {}
    for key, value in dic.items():
        target = key_mapping.get(key, key)
        for t in ([target] if isinstance(target, str) else target):
            out[t] = value
    return out


def insertAfter(lst, key, item):
    for pos, entry in enumerate(lst):
        if entry[0] == key:
            return list(lst[:pos + 1]) + [item] + list(lst[pos + 1:])
    raise ValueError(f"Key {key} not found.")


def replace(lst, key, item):
    for pos, entry in enumerate(lst):
        if entry[0] == key:
            return list(lst[:pos]) + [item] + list(lst[pos + 1:])
    raise ValueError(f"Key {key} not found.")


def _countParameters(module):
    return sum(p.numel() for p in module.parameters() if p.requires_grad)


def getScaler(operations):
    """batch -> batch with per-key scale / shift ops, e.g. [('CA', ('shift', 'mean')), ('CA', ('scale', 1/25.83))]
    (reference utils.py:15-47); segment sums run on the batch's device."""
    def scaler(batch):
        batch = batch.clone()
        seg = batch.nodeSegment()
        for key, op in operations:
            keys = key if isinstance(key, (tuple, list)) else [key]
            if op[0] == "scale":
                for k in keys:
                    batch[k] = batch[k] * op[1]
            elif op[0] == "shift" and op[1] == "mean":
                n = batch["_n_nodes"].view(-1, 1).to(batch[key].dtype)
                tot = torch.zeros(n.shape[0], batch[key].shape[1], dtype=batch[key].dtype, device=batch[key].device)
                tot.index_add_(0, seg.to(batch[key].device), batch[key])
                batch[key] = batch[key] - (tot / n)[seg]
            elif op[0] == "shift" and op[1] in batch:
                sign = op[2] if len(op) == 3 else 1
                batch[key] = batch[key] + sign * batch[op[1]]
            else:
                raise ValueError(op)
        return batch
    return scaler
