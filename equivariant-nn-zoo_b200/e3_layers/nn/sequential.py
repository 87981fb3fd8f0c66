"""Layer protocol and the sequential graph network (reference ``e3_layers/nn/sequential.py``).

Every layer is ``forward(data: dict, attrs: dict) -> (new_data, new_attrs)``; irreps arguments
are ``irreps | (irreps, custom_key)`` and the runner renames keys on the way in and out."""
from collections import OrderedDict

import torch
from torch.profiler import record_function

from e3b200.irreps import Irreps

from ..data import Batch
from ..utils import ConfigDict, build, keyMap


class Module(torch.nn.Module):
    def init_irreps(self, output_keys=(), **kwargs):
        outs = {output_keys} if isinstance(output_keys, str) else set(output_keys)
        self.irreps_in, self.irreps_out = {}, {}
        self.input_key_mapping, self.output_key_mapping = {}, {}
        for name, spec in kwargs.items():
            if spec is None:
                continue
            if isinstance(spec, (str, Irreps)):
                irreps, custom = spec, name
            else:
                irreps, custom = spec
            if name in outs:
                self.irreps_out[name] = irreps
                self.output_key_mapping[name] = custom
            else:
                self.irreps_in[name] = irreps
                self.input_key_mapping[custom] = name

    def inputKeyMap(self, obj):
        return keyMap(obj, self.input_key_mapping)

    def outputKeyMap(self, obj):
        return keyMap(obj, self.output_key_mapping)


def strip_reference_state(state, own_keys=None):
    """A checkpoint written by the reference (e3nn 0.4.4 modules under DistributedDataParallel) carries keys this
    implementation has no use for: the ``module.`` prefix of DDP (reference ``inference.py:46-53`` strips it too), the
    Wigner-3j constants and generated sub-modules of e3nn's compiled tensor products / linear maps
    (``..._compiled_main_left_right._w3j_1_1_0`` etc.), and buffers of e3nn layers that are pure functions here.  They
    are dropped; every parameter and every buffer this model owns must still be present (strict loading)."""
    out = {}
    for k, v in state.items():
        k = k[7:] if k.startswith("module.") else k
        if "_compiled_main" in k or "._w3j_" in k or k.rsplit(".", 1)[-1].startswith("_w3j_"):
            continue
        if own_keys is not None and k not in own_keys and k.rsplit(".", 1)[-1] in ("output_mask", "_zeros", "cst"):
            continue
        out[k] = v
    return out


class SequentialGraphNetwork(torch.nn.Sequential):
    """Runs (key, layer) pairs over one shared dict; a layer is a ``Module`` built from a config
    node or any callable ``f(data, attrs)``.  ``jit`` in the config is accepted and ignored (the
    kernels are already compiled)."""

    def __init__(self, **config):
        built, self.layers = OrderedDict(), []
        for key, node in config["layers"]:
            if isinstance(node, (dict, ConfigDict)):
                layer = build(node)
                built[key] = layer
            elif callable(node):
                layer = node
            else:
                raise TypeError("invalid config node")
            self.layers.append((key, layer))
        self.layer_configs = config["layers"]
        super().__init__(built)
        # the interaction blocks of one network share work that depends on the edges only (their radial hidden layers run
        # as one grouped launch per layer, e3b200.interaction._RadialHidden)
        blocks = [m for m in built.values() if type(m).__name__ == "MessagePassing"]
        for m in blocks:
            object.__setattr__(m, "_e3b_group", blocks)

    def load_state_dict(self, state_dict, strict=True, **kwargs):
        """accepts the reference's checkpoints as they are (see ``strip_reference_state``)"""
        return super().load_state_dict(strip_reference_state(state_dict, set(self.state_dict().keys())), strict=strict, **kwargs)

    def forward(self, batch):
        data, attrs = batch.data, batch.attrs
        for key, layer in self.layers:
            with record_function(key):
                mapped = isinstance(layer, Module)
                d = layer.inputKeyMap(data) if mapped else data
                a = layer.inputKeyMap(attrs) if mapped else attrs
                d, a = layer(d, a)
                if mapped:
                    d, a = layer.outputKeyMap(d), layer.outputKeyMap(a)
                data.update(d)
                attrs.update(a)
        return Batch(attrs, **data)
