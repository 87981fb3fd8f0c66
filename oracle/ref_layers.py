"""CPU restatement of the reference's hot-path layers in the REFERENCE dataflow
(oracle; test infrastructure only).

Every class cites the reference file:line it follows.  Differences from the product are
deliberate: per-edge materialisation of gathered features / radial weights / tensor-product
output, the per-edge ``o3.Linear`` BEFORE the scatter (``nn/pointwise.py:99``), the dense
all-pairs neighbour list, ``scatter_add`` reductions and autograd for forces.

Layers follow the reference protocol ``forward(data, attrs) -> (new_data, new_attrs)`` on plain
dicts; ``OracleNetwork`` replays ``SequentialGraphNetwork.forward`` (``nn/sequential.py:70-88``).
The model is built from a config-node list ``[(key, {"module": cls, **kwargs} | callable)]``;
classes/callables are resolved BY NAME so the same node list can come from the genuine
reference configs, from the product's mirror of them, or be written by hand.
"""
import functools
import math
from typing import Dict

import torch
from torch import nn

from . import e3nn_ops as e3
from .irreps import Irrep, Irreps

# ---- activations: e3_layers/utils/utils.py:69-84 ------------------------------------------


def tanhlu(x):
    return torch.tanh(x) * torch.abs(x)


def ShiftedSoftPlus(x):
    return torch.nn.functional.softplus(x) - math.log(2.0)


activations = {
    "abs": torch.abs,
    "tanh": torch.tanh,
    "ssp": ShiftedSoftPlus,
    "silu": torch.nn.functional.silu,
    "tanhlu": tanhlu,
}


def tp_path_exists(irreps_in1, irreps_in2, ir_out):
    """utils/utils.py:87-96"""
    irreps_in1 = Irreps(irreps_in1).simplify()
    irreps_in2 = Irreps(irreps_in2).simplify()
    ir_out = Irrep(ir_out)
    return any(ir_out in ir1 * ir2 for _, ir1 in irreps_in1 for _, ir2 in irreps_in2)


def key_map(dic, mapping):
    """utils/utils.py:139-152 (dict branch)"""
    out = {}
    for k, v in dic.items():
        if k in mapping:
            nk = mapping[k]
            if isinstance(nk, str):
                out[nk] = v
            else:
                for item in nk:
                    out[item] = v
        else:
            out[k] = v
    return out


class Module(nn.Module):
    """nn/sequential.py:12-39"""

    def init_irreps(self, output_keys=(), **kwargs):
        if isinstance(output_keys, str):
            output_keys = [output_keys]
        self.irreps_in, self.irreps_out = {}, {}
        self.input_key_mapping, self.output_key_mapping = {}, {}
        for key, value in kwargs.items():
            if isinstance(value, (str, Irreps)):
                irreps, custom = value, key
            elif isinstance(value, (list, tuple)):
                assert len(value) == 2
                irreps, custom = value
            else:
                continue
            if key in output_keys:
                self.irreps_out[key] = irreps
                self.output_key_mapping[key] = custom
            else:
                self.irreps_in[key] = irreps
                self.input_key_mapping[custom] = key


# ---- neighbour list / edge vectors: e3_layers/data/compute_edge.py -------------------------


def computeEdgeVector(data, attrs, key="pos", with_lengths=True):
    """data/compute_edge.py:13-36: edge_vector = pos[ei[1]] - pos[ei[0]]."""
    attrs["edge_vector"] = ("edge", "1x1o")
    attrs["edge_length"] = ("edge", "1x0e")
    if "edge_vector" in data:
        if with_lengths and "edge_length" not in data:
            data["edge_length"] = torch.linalg.norm(data["edge_vector"], dim=-1)
        return data, attrs
    pos, ei = data[key], data["edge_index"]
    vec = pos[ei[1]] - pos[ei[0]]
    data["edge_vector"] = vec
    if with_lengths:
        data["edge_length"] = torch.linalg.norm(vec, dim=-1)
    return data, attrs


def _compute_edge_map(a, b):
    """computeEdgeMap, data/compute_edge.py:77-84: position in b of every column of a (both sorted alike)"""
    j, lst = 0, []
    for i in range(a.shape[1]):
        while not bool((b[:, j] == a[:, i]).all()):
            j += 1
        lst.append(j)
    return torch.tensor(lst, dtype=torch.long)


def computeEdgeIndex(data, attrs, r_max=None, key="pos", criteria=None):
    """data/compute_edge.py:38-113: per graph all ordered pairs (a slow, b fast), keep ||pos[a]-pos[b]|| < r_max
    (fp32 norm, strict) OR criteria, AND a != b; then (:86-100) a pre-existing edge list is re-added and the per-edge
    tensors are mapped to the new numbering, zero-padded for the new edges.  Returns only the new keys (:110-113)."""
    pos = torch.as_tensor(data[key], dtype=torch.get_default_dtype())
    n_nodes = data["_n_nodes"].reshape(-1).tolist()
    chunks, cnt = [], 0
    for n in n_nodes:
        ar = torch.arange(cnt, cnt + n)
        a = ar.repeat_interleave(n)
        b = ar.repeat(n)
        chunks.append(torch.stack([a, b]))
        cnt += n
    ei = torch.cat(chunks, dim=1) if chunks else torch.zeros(2, 0, dtype=torch.long)
    dist = torch.linalg.norm(pos[ei[0]] - pos[ei[1]], dim=-1)
    mask = dist < r_max
    if criteria is not None:
        mask = torch.logical_or(mask, criteria(data, ei))
    mask = torch.logical_and(mask, ei[0] != ei[1])
    if "edge_index" in data:                                       # :86-88
        mask[_compute_edge_map(data["edge_index"], ei)] = True
    ei = ei[:, mask]
    if "edge_index" in data:                                       # :93-100
        edge_map = _compute_edge_map(data["edge_index"], ei)
        for k in attrs:
            if attrs[k][0] == "edge":
                tmp = data[k]
                data[k] = torch.zeros(ei.shape[1], tmp.shape[1], dtype=tmp.dtype)
                data[k][edge_map] = tmp
    seg = torch.repeat_interleave(torch.arange(len(n_nodes)), torch.tensor(n_nodes, dtype=torch.long))
    n_edges = torch.bincount(seg[ei[0]], minlength=len(n_nodes)).view(-1, 1)
    attrs["_n_edges"] = ("graph", "1x0e")
    data["_n_edges"] = n_edges
    return {"edge_index": ei}, attrs


# ---- embeddings: e3_layers/nn/embedding.py ---------------------------------------------------


def symmetricCutoff(x, factor, p=6.0):
    """nn/embedding.py:26-29"""
    x = x * factor
    return (x - 1) ** 2 * (x + 1) ** 2 * (abs(x) < 1.0).float()


def _poly_cutoff(x, factor, p=6.0):
    """nn/embedding.py:31-40"""
    x = x * factor
    out = 1.0
    out = out - (((p + 1.0) * (p + 2.0) / 2.0) * torch.pow(x, p))
    out = out + (p * (p + 2.0) * torch.pow(x, p + 1.0))
    out = out - ((p * (p + 1.0) / 2) * torch.pow(x, p + 2.0))
    return out * (x < 1.0)


class BesselBasis(nn.Module):
    """nn/embedding.py:74-127"""

    def __init__(self, r_max, r_min=0, num_basis=8, trainable=True, one_over_r=True):
        super().__init__()
        self.r_max, self.r_min = float(r_max), float(r_min)
        self.prefactor = 2.0 / (self.r_max - self.r_min)
        self.one_over_r = one_over_r
        w = torch.linspace(start=1.0, end=num_basis, steps=num_basis) * math.pi
        if trainable:
            self.bessel_weights = nn.Parameter(w)
        else:
            self.register_buffer("bessel_weights", w)

    def forward(self, x):
        num = torch.sin(self.bessel_weights * x.unsqueeze(-1) / (self.r_max - self.r_min))
        res = self.prefactor * num
        if self.one_over_r:
            res = res / x.unsqueeze(-1)
        return res


class PolynomialCutoff(nn.Module):
    """nn/embedding.py:43-71"""

    def __init__(self, r_max, p=6, cutoff=_poly_cutoff):
        super().__init__()
        self.p = float(p)
        self._factor = 1.0 / float(r_max)
        self.cutoff = cutoff

    def forward(self, x):
        return self.cutoff(x, self._factor, p=self.p)


class SphericalEncoding(Module):
    """nn/embedding.py:130-178"""

    def __init__(self, irreps_out, edge_sh_normalization="component", edge_sh_normalize=True, irreps_in="1x1o"):
        super().__init__()
        self.init_irreps(vectors=irreps_in, spherical_harmonics=irreps_out, output_keys=["spherical_harmonics"])
        self.mul = Irreps(self.irreps_in["vectors"])[0].mul
        irr = []
        for mi in Irreps(self.irreps_out["spherical_harmonics"]):
            assert mi.mul == self.mul
            irr.append(str(mi.ir))
        self.sh = e3.SphericalHarmonics("+".join(irr), edge_sh_normalize, edge_sh_normalization)

    def forward(self, data, attrs):
        v = data["vectors"]
        cat = v.shape[0]
        sh = self.sh(v.view(cat, self.mul, 3)).view(cat, -1)
        return ({"spherical_harmonics": sh},
                {"spherical_harmonics": ("edge", self.irreps_out["spherical_harmonics"])})


class RadialBasisEncoding(Module):
    """nn/embedding.py:181-219"""

    def __init__(self, r_max, trainable, irreps_out, r_min=0, polynomial_degree=6, basis=None,
                 cutoff=None, irreps_in="1x0e", one_over_r=True):
        super().__init__()
        self.init_irreps(input=irreps_in, radial_embedding=irreps_out, output_keys=["radial_embedding"])
        num_basis = Irreps(self.irreps_out["radial_embedding"])[0].mul
        cutoff_fn = _poly_cutoff
        if cutoff is not None and (cutoff == "symmetricCutoff" or getattr(cutoff, "__name__", "") == "symmetricCutoff"):
            cutoff_fn = symmetricCutoff
        self.basis = BesselBasis(r_max, r_min, num_basis, trainable, one_over_r=one_over_r)
        self.cutoff = PolynomialCutoff(r_max, p=polynomial_degree, cutoff=cutoff_fn)
        self.r_max = r_max

    def forward(self, data, attrs):
        x = data["input"]
        emb = (self.basis(x) * self.cutoff(x)[:, None]).view(x.shape[0], -1)
        return ({"radial_embedding": emb},
                {"radial_embedding": (attrs["input"][0], self.irreps_out["radial_embedding"])})


class Broadcast(Module):
    """nn/embedding.py:222-254"""

    def __init__(self, irreps_in, irreps_out, to):
        super().__init__()
        self.init_irreps(input=irreps_in, output=irreps_out, output_keys=["output"])
        self.to = to

    def forward(self, data, attrs):
        assert attrs["input"][0] == "graph"
        seg = data["_node_segment"] if self.to == "node" else data["_edge_segment"]
        return {"output": data["input"][seg]}, {"output": (self.to, self.irreps_out["output"])}


class OneHotEncoding(Module):
    """nn/embedding.py:257-281"""

    def __init__(self, num_types, irreps_out, irreps_in="0x0e"):
        super().__init__()
        self.num_types = num_types
        self.init_irreps(input=irreps_in, one_hot=irreps_out, output_keys="one_hot")

    def forward(self, data, attrs):
        t = data["input"].squeeze(-1)
        oh = torch.nn.functional.one_hot(t, num_classes=self.num_types).to(torch.get_default_dtype())
        return {"one_hot": oh}, {"one_hot": (attrs["input"][0], self.irreps_out["one_hot"])}


class RelativePositionEncoding(Module):
    """nn/embedding.py:283-312"""

    def __init__(self, radial_encoding, segment, irreps_out, id=None):
        super().__init__()
        self.init_irreps(input=segment, output=irreps_out, id=id, output_keys=["output"])
        radial_encoding = dict(radial_encoding)
        radial_encoding["irreps_in"] = "1x0e"
        radial_encoding["irreps_out"] = self.irreps_out["output"]
        self.radial = build(radial_encoding)

    def forward(self, data, attrs):
        seg, ei = data["input"], data["edge_index"]
        if "id" in self.irreps_in and self.irreps_in["id"] is not None:
            idt = data["id"]
            rel = idt[ei[0]] - idt[ei[1]]
        else:
            rel = ei[0] - ei[1]
        mask = (seg[ei[0]] == seg[ei[1]]).float()  # literally float32, as the reference (:308)
        rel = mask * rel.view(-1, 1) + (1 - mask) * 1e5
        out, _ = self.radial({"input": rel}, attrs)
        return {"output": out["radial_embedding"]}, {"output": ("edge", self.irreps_out["output"])}


# ---- pointwise: e3_layers/nn/pointwise.py ---------------------------------------------------


class PointwiseLinear(Module):
    """nn/pointwise.py:14-30"""

    def __init__(self, irreps_in, irreps_out, biases=True, **kwargs):
        super().__init__()
        self.init_irreps(input=irreps_in, output=irreps_out, output_keys=["output"])
        self.linear = e3.Linear(self.irreps_in["input"], self.irreps_out["output"], biases=biases)

    def forward(self, data, attrs):
        return ({"output": self.linear(data["input"])},
                {"output": (attrs["input"][0], self.irreps_out["output"])})


class LayerNormalization(Module):
    """nn/pointwise.py:32-51"""

    def __init__(self, irreps_in, irreps_out, **kwargs):
        super().__init__()
        self.init_irreps(input=irreps_in, output=irreps_out, output_keys=["output"])
        assert irreps_in == irreps_out
        self.muls = [mi.mul for mi in Irreps(irreps_in)]
        self.slices = [(s.start, s.stop) for s in Irreps(irreps_in).slices()]
        self.std = nn.Parameter(torch.ones(len(self.slices)))

    def forward(self, data, attrs):
        x = data["input"]
        out = torch.zeros_like(x)
        for i, (a, b) in enumerate(self.slices):
            t = x[:, a:b]
            norm = ((t * t).sum(dim=-1, keepdim=True) / self.muls[i] + 1e-6) ** 0.5
            out[:, a:b] = t / norm * self.std[i]
        return {"output": out}, attrs


class TensorProductExpansion(Module):
    """nn/pointwise.py:54-100: instruction builder + o3.TensorProduct + post-Linear."""

    def __init__(self, left, right, output, instruction="uvu", internal_weight=True, **kwargs):
        super().__init__()
        self.init_irreps(left=left, right=right, output=output, output_keys=["output"])
        mid, instr = [], []
        for i, (mul, irl) in enumerate(Irreps(self.irreps_in["left"])):
            for j, (_, irr) in enumerate(Irreps(self.irreps_in["right"])):
                for ir_out in irl * irr:
                    if ir_out in Irreps(self.irreps_out["output"]):
                        k = len(mid)
                        mid.append((mul, ir_out))
                        instr.append((i, j, k, instruction, True))
        mid, p, _ = Irreps(mid).sort()
        instr = [(a, b, p[c], m, t) for a, b, c, m, t in instr]
        self.tp = e3.TensorProduct(Irreps(self.irreps_in["left"]), Irreps(self.irreps_in["right"]), mid, instr,
                                   shared_weights=internal_weight, internal_weights=internal_weight)
        self.internal_weight = internal_weight
        self.linear = e3.Linear(mid.simplify(), Irreps(self.irreps_out["output"]))

    def forward(self, left=None, right=None, weight=None):
        out = self.tp(left, right) if self.internal_weight else self.tp(left, right, weight)
        return self.linear(out)


class Concat(Module):
    """nn/pointwise.py:134-152"""

    def __init__(self, irreps_out, **irreps_in):
        super().__init__()
        self.init_irreps(**irreps_in, output=irreps_out, output_keys=["output"])
        lst = [Irreps(v) for v in self.irreps_in.values()]
        tot = lst[0]
        for x in lst[1:]:
            tot = tot + x
        self.linear = e3.Linear(tot, Irreps(self.irreps_out["output"]), biases=True)

    def forward(self, data, attrs):
        x = torch.cat([data[k] for k in self.irreps_in.keys()], dim=1)
        key = list(self.irreps_in.keys())[0]
        return {"output": self.linear(x)}, {"output": (attrs[key][0], self.irreps_out["output"])}


# ---- message passing: e3_layers/nn/message_passing.py ---------------------------------------


class FactorizedConvolution(Module):
    """nn/message_passing.py:21-124"""

    def __init__(self, input_features, output_features, node_attrs, edge_radial, edge_spherical,
                 invariant_layers=1, invariant_neurons=8, avg_num_neighbors=None, use_sc=True,
                 nonlinearity_scalars=None, reduce=True):
        super().__init__()
        self.init_irreps(input_features=input_features, output_features=output_features,
                         node_attrs=node_attrs, edge_radial=edge_radial, edge_spherical=edge_spherical,
                         output_keys=["output_features"])
        self.avg_num_neighbors, self.use_sc = avg_num_neighbors, use_sc
        fin = self.irreps_in["input_features"]
        fout = self.irreps_out["output_features"]
        self.linear_1 = e3.Linear(fin, fin)
        self.tp = TensorProductExpansion(fin, (self.irreps_in["edge_spherical"], "edge_spherical"),
                                         (fout, "edge_features"), "uvu", internal_weight=False)
        self.fc = e3.FullyConnectedNet(
            [Irreps(self.irreps_in["edge_radial"]).num_irreps] + invariant_layers * [invariant_neurons]
            + [self.tp.tp.weight_numel], activations["ssp"])
        self.sc = None
        if use_sc:
            self.sc = e3.FullyConnectedTensorProduct(fin, Irreps(self.irreps_in["node_attrs"]), fout)
        self.reduce = reduce

    def forward(self, data, attrs):
        weight = self.fc(data["edge_radial"])
        x = data["input_features"]
        src, dst = data["edge_index"][0], data["edge_index"][1]
        sc = self.sc(x, data["node_attrs"]) if self.sc is not None else None
        x = self.linear_1(x)
        ef = self.tp(left=x[src], right=data["edge_spherical"], weight=weight)
        if self.reduce:
            x = e3.scatter(ef, dst, dim=0, dim_size=len(x))
            if self.avg_num_neighbors is not None:
                x = x.div(self.avg_num_neighbors ** 0.5)
            if sc is not None:
                x = x + sc
        else:
            x = ef
        return ({"output_features": x},
                {"output_features": (attrs["input_features"][0], self.irreps_out["output_features"])})


class MessagePassing(Module):
    """nn/message_passing.py:127-262 (gate nonlinearity only; 'norm' is never selected)."""

    def __init__(self, input_features, output_features, node_attrs, edge_radial, edge_spherical,
                 convolution, resnet=False, nonlinearity_type="gate",
                 nonlinearity_scalars=None, nonlinearity_gates=None, normalize=False):
        super().__init__()
        nonlinearity_scalars = nonlinearity_scalars or {"e": "ssp", "o": "tanh"}
        nonlinearity_gates = nonlinearity_gates or {"e": "ssp", "o": "abs"}
        self.init_irreps(input_features=input_features, output_features=output_features,
                         node_attrs=node_attrs, edge_radial=edge_radial, edge_spherical=edge_spherical,
                         output_keys=["output_features"])
        assert nonlinearity_type == "gate"
        ns = {1: nonlinearity_scalars["e"], -1: nonlinearity_scalars["o"]}
        ng = {1: nonlinearity_gates["e"], -1: nonlinearity_gates["o"]}
        sh = Irreps(self.irreps_in["edge_spherical"])
        prev = Irreps(self.irreps_in["input_features"])
        hidden = Irreps(self.irreps_out["output_features"])
        scalars = Irreps([(m, ir) for m, ir in hidden if ir.l == 0 and tp_path_exists(prev, sh, ir)])
        gated = Irreps([(m, ir) for m, ir in hidden if ir.l > 0 and tp_path_exists(prev, sh, ir)])
        layer_out = (scalars + gated).simplify()
        gates = Irreps([(m, "0e") for m, _ in gated])
        self.equivariant_nonlin = e3.Gate(
            scalars, [activations[ns[ir.p]] for _, ir in scalars],
            gates, [activations[ng[ir.p]] for _, ir in gates], gated)
        conv_out = self.equivariant_nonlin.irreps_in.simplify()
        self.resnet = bool(layer_out == prev and resnet)
        conv = dict(convolution)
        self.conv = build(conv, input_features=input_features, output_features=conv_out,
                          node_attrs=node_attrs, edge_radial=edge_radial, edge_spherical=edge_spherical)
        self.normalize = normalize
        if normalize:
            self.norm = LayerNormalization(self.irreps_out["output_features"], self.irreps_out["output_features"])

    def forward(self, data, attrs):
        old = data["input_features"]
        d, _ = self.conv(data, attrs)
        out = self.equivariant_nonlin(d["output_features"])
        if self.resnet:
            out = old + out
        if self.normalize:
            out = self.norm({"input": out}, attrs)[0]["output"]
        return ({"output_features": out},
                {"output_features": (attrs["input_features"][0], self.irreps_out["output_features"])})


# ---- heads: e3_layers/nn/scaling.py, e3_layers/nn/output.py --------------------------------


class PerTypeScaleShift(Module):
    """nn/scaling.py:9-67"""

    def __init__(self, num_types, shifts, scales, scales_trainable=False, shifts_trainable=False,
                 irreps_in="1x0e", irreps_out="1x0e", species="1x0e"):
        super().__init__()
        self.init_irreps(input=irreps_in, output=irreps_out, species=species, output_keys=["output"])
        self.has_shifts, self.has_scales = shifts is not None, scales is not None
        for name, val, tr in (("shifts", shifts, shifts_trainable), ("scales", scales, scales_trainable)):
            if val is None:
                continue
            t = torch.as_tensor(val, dtype=torch.get_default_dtype())
            if t.numel() == 1:
                t = torch.ones(num_types) * t
            assert t.shape == (num_types,)
            if tr:
                setattr(self, name, nn.Parameter(t))
            else:
                self.register_buffer(name, t)

    def forward(self, data, attrs):
        sp, x = data["species"], data["input"]
        if self.has_scales:
            x = self.scales[sp].view(-1, 1) * x
        if self.has_shifts:
            x = self.shifts[sp].view(-1, 1) + x
        return {"output": x}, {"output": (attrs["input"][0], self.irreps_out["output"])}


class Pooling(Module):
    """nn/output.py:56-74"""

    def __init__(self, irreps_in, irreps_out, reduce):
        super().__init__()
        self.init_irreps(input=irreps_in, output=irreps_out, output_keys=["output"])
        assert reduce == "sum"

    def forward(self, data, attrs):
        out = e3.scatter(data["input"], data["_node_segment"], dim=0)
        return {"output": out}, {"output": ("graph", self.irreps_out["output"])}


class OracleNetwork(nn.Module):
    """nn/sequential.py:42-88 on plain dicts."""

    def __init__(self, layers, **_ignored):
        super().__init__()
        self.layers = []
        for key, value in layers:
            if isinstance(value, dict) or hasattr(value, "to_dict"):
                m = build(value)
                self.add_module(key, m)
                self.layers.append((key, m))
            elif callable(value) or isinstance(value, (str, tuple)):
                self.layers.append((key, _resolve_callable(value)))
            else:
                raise TypeError("invalid config node")

    def forward(self, data, attrs):
        data, attrs = dict(data), dict(attrs)
        if "_n_nodes" in data and "_node_segment" not in data:
            n = data["_n_nodes"].reshape(-1)
            data["_node_segment"] = torch.repeat_interleave(torch.arange(len(n)), n)
        for key, m in self.layers:
            if "_n_edges" in data and "_edge_segment" not in data:
                n = data["_n_edges"].reshape(-1)
                data["_edge_segment"] = torch.repeat_interleave(torch.arange(len(n)), n)
            d, a = data, attrs
            if isinstance(m, Module):
                d, a = key_map(d, m.input_key_mapping), key_map(a, m.input_key_mapping)
            d, a = m(d, a)
            if isinstance(m, Module):
                d, a = key_map(d, m.output_key_mapping), key_map(a, m.output_key_mapping)
            data.update(d)
            attrs.update(a)
        return data, attrs


SequentialGraphNetwork = OracleNetwork


class GradientOutput(Module):
    """nn/output.py:18-53: gradients = sign * d(sum y)/dx via autograd."""

    def __init__(self, func, x, y, gradients, sign=1.0, **kwargs):
        super().__init__()
        self.sign = float(sign)
        assert self.sign in (1.0, -1.0)
        self.init_irreps(x=x, y=y, gradients=gradients, output_keys=["gradients"])
        if isinstance(func, dict) or hasattr(func, "to_dict"):
            func = build(func, **kwargs)
        self.func = func

    def forward(self, data, attrs, create_graph=False):
        data = dict(data)
        xkey = [k for k, v in self.input_key_mapping.items() if v == "x"][0]
        ykey = [k for k, v in self.input_key_mapping.items() if v == "y"][0]
        x = data[xkey].detach().clone().requires_grad_(True)
        data[xkey] = x
        out, oattrs = self.func(data, attrs)
        (g,) = torch.autograd.grad(out[ykey].sum(), x, create_graph=create_graph)
        gkey = self.output_key_mapping["gradients"]
        out[gkey] = self.sign * g
        oattrs[gkey] = (attrs[xkey][0], self.irreps_out["gradients"])
        return out, oattrs


_BY_NAME = {c.__name__: c for c in (
    SphericalEncoding, RadialBasisEncoding, Broadcast, OneHotEncoding, RelativePositionEncoding,
    PointwiseLinear, LayerNormalization, TensorProductExpansion, Concat, FactorizedConvolution,
    MessagePassing, PerTypeScaleShift, Pooling, GradientOutput)}
_BY_NAME["SequentialGraphNetwork"] = OracleNetwork
_FUNCS = {"computeEdgeVector": computeEdgeVector, "computeEdgeIndex": computeEdgeIndex}


def _resolve_callable(fn):
    if isinstance(fn, str):
        return _FUNCS[fn]
    if isinstance(fn, tuple):
        return functools.partial(_FUNCS[fn[0]], **fn[1])
    if isinstance(fn, functools.partial):
        base = _FUNCS[fn.func.__name__]
        return functools.partial(base, *fn.args, **fn.keywords)
    return _FUNCS[fn.__name__]


def _plain(node):
    return node.to_dict() if hasattr(node, "to_dict") else dict(node)


def build(node, **kwargs):
    """utils/utils.py:99-116 with classes resolved by name into this module."""
    import inspect

    node = _plain(node)
    cls = node["module"]
    cls = _BY_NAME[cls if isinstance(cls, str) else cls.__name__]
    kwargs.update(node)
    kwargs.pop("module")
    spec = inspect.getfullargspec(cls.__init__)
    if not spec.varkw:
        names = inspect.signature(cls.__init__).parameters
        kwargs = {k: v for k, v in kwargs.items() if k in names}
    return cls(**kwargs)
