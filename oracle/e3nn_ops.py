"""e3nn 0.4.4 operators used by the reference's hot path, restated in plain torch
(oracle; test infrastructure only; CPU, fp32 or fp64 via torch default dtype).

Each class names the reference call site it serves and the SURVEY.md appendix item that
restates the e3nn semantics it follows (e3nn itself is an un-vendored third-party dependency
pinned at 0.4.4 in ``/root/reference/requirements.txt:27``).
"""
import math
from typing import List, Optional

import torch
from torch import nn

from .irreps import Irrep, Irreps
from .wigner import spherical_harmonics, wigner_3j

# ------------------------------------------------------------------------------------------
# normalize2mom.  c = E_{z~N(0,1)}[act(z)^2]^(-1/2), estimated by e3nn with
# 1e6 float64 samples from torch.Generator('cpu').manual_seed(0).  The table holds the values
# that recipe yields with torch 2.11 CPU; `normalize2mom_constant` recomputes them.
NORMALIZE2MOM = {
    "ssp": 1.878204668541552,
    "silu": 1.6791767923989418,
    "tanh": 1.5937334472592692,
    "tanhlu": 1.1501944455736173,
    "abs": 1.001110600838467,
}


_N2M_CACHE = {}


def normalize2mom_constant(f):
    key = id(f)
    if key in _N2M_CACHE and _N2M_CACHE[key][0] is f:
        return _N2M_CACHE[key][1]
    c = _normalize2mom_constant(f)
    _N2M_CACHE[key] = (f, c)
    return c


def _normalize2mom_constant(f):
    gen = torch.Generator(device="cpu").manual_seed(0)
    z = torch.randn(1_000_000, generator=gen, dtype=torch.float64)
    with torch.no_grad():
        return f(z).pow(2).mean().pow(-0.5).item()


class normalize2mom(nn.Module):
    """e3nn.math.normalize2mom: wraps an activation as c*act(x); identity scale if |c-1|<1e-4."""

    def __init__(self, f):
        super().__init__()
        cst = normalize2mom_constant(f)
        self._is_id = abs(cst - 1) < 1e-4
        self.f = f
        self.cst = cst

    def forward(self, x):
        if self._is_id:
            return self.f(x)
        return self.f(x).mul(self.cst)


# ------------------------------------------------------------------------------------------
class Linear(nn.Module):
    """o3.Linear.  Reference call sites: ``nn/message_passing.py:58-63`` (linear_1),
    ``nn/pointwise.py:18-22`` (PointwiseLinear), ``:87-92`` (tp.linear), ``:142`` (Concat)."""

    def __init__(self, irreps_in, irreps_out, internal_weights=True, shared_weights=True, biases=False):
        super().__init__()
        self.irreps_in = Irreps(irreps_in)
        self.irreps_out = Irreps(irreps_out)
        assert internal_weights and shared_weights
        ins = []
        for i_in, (mul_in, ir_in) in enumerate(self.irreps_in):
            for i_out, (mul_out, ir_out) in enumerate(self.irreps_out):
                if ir_in == ir_out:
                    ins.append((i_in, i_out, mul_in, mul_out))
        fan = {}
        for i_in, i_out, mul_in, mul_out in ins:
            fan[i_out] = fan.get(i_out, 0) + mul_in
        self.instructions = [
            (i_in, i_out, mul_in, mul_out, (fan[i_out] if fan[i_out] > 0 else 1.0) ** -0.5)
            for i_in, i_out, mul_in, mul_out in ins
        ]
        self.weight_numel = sum(mi * mo for _, _, mi, mo, _ in self.instructions)
        self.weight = nn.Parameter(torch.randn(self.weight_numel))
        self.bias_blocks = [
            i for i, (mul, ir) in enumerate(self.irreps_out) if biases and ir.is_scalar()
        ]
        self.bias_numel = sum(self.irreps_out[i].mul for i in self.bias_blocks)
        if self.bias_numel > 0:
            self.bias = nn.Parameter(torch.zeros(self.bias_numel))
        else:
            self.register_buffer("bias", torch.zeros(0))
        touched = {i_out for _, i_out, _, _, _ in self.instructions} | set(self.bias_blocks)
        mask = torch.cat(
            [
                (torch.ones if i in touched else torch.zeros)(mul * ir.dim)
                for i, (mul, ir) in enumerate(self.irreps_out)
            ]
        ) if len(self.irreps_out) else torch.zeros(0)
        self.register_buffer("output_mask", mask)

    def forward(self, x):
        lead = x.shape[:-1]
        x = x.reshape(-1, self.irreps_in.dim)
        z = x.shape[0]
        sl_in = self.irreps_in.slices()
        outs = [None] * len(self.irreps_out)
        off = 0
        for i_in, i_out, mul_in, mul_out, alpha in self.instructions:
            w = self.weight[off : off + mul_in * mul_out].reshape(mul_in, mul_out)
            off += mul_in * mul_out
            d = self.irreps_in[i_in].ir.dim
            xb = x[:, sl_in[i_in]].reshape(z, mul_in, d)
            y = alpha * torch.einsum("uw,zui->zwi", w, xb)
            outs[i_out] = y if outs[i_out] is None else outs[i_out] + y
        boff = 0
        for i in self.bias_blocks:
            mul = self.irreps_out[i].mul
            b = self.bias[boff : boff + mul].reshape(1, mul, 1)
            boff += mul
            outs[i] = b.expand(z, mul, 1) if outs[i] is None else outs[i] + b
        cols = []
        for i, (mul, ir) in enumerate(self.irreps_out):
            if outs[i] is None:
                cols.append(x.new_zeros(z, mul * ir.dim))
            else:
                cols.append(outs[i].reshape(z, mul * ir.dim))
        out = torch.cat(cols, dim=1) if cols else x.new_zeros(z, 0)
        return out.reshape(*lead, self.irreps_out.dim)


# ------------------------------------------------------------------------------------------
class TensorProduct(nn.Module):
    """o3.TensorProduct, modes 'uvu', 'uvw', 'uuu'; component / element
    normalisation.  Reference call site: ``nn/pointwise.py:78-85`` (uvu, external weights)."""

    def __init__(self, irreps_in1, irreps_in2, irreps_out, instructions,
                 shared_weights=None, internal_weights=None):
        super().__init__()
        self.irreps_in1 = Irreps(irreps_in1)
        self.irreps_in2 = Irreps(irreps_in2)
        self.irreps_out = Irreps(irreps_out)
        norm_ins = []
        for ins in instructions:
            ins = tuple(ins)
            if len(ins) == 5:
                ins = ins + (1.0,)
            norm_ins.append(ins)

        def num_elements(i1, i2, mode):
            m1, m2 = self.irreps_in1[i1].mul, self.irreps_in2[i2].mul
            return {"uvw": m1 * m2, "uvu": m2, "uvv": m1, "uuw": m1, "uuu": 1, "uvuv": 1}[mode]

        self.instructions = []
        for (i1, i2, io, mode, has_w, pw) in norm_ins:
            m1, ir1 = self.irreps_in1[i1]
            m2, ir2 = self.irreps_in2[i2]
            mo, iro = self.irreps_out[io]
            assert iro in ir1 * ir2
            alpha = iro.dim
            x = sum(num_elements(j1, j2, md) for (j1, j2, jo, md, _, _) in norm_ins if jo == io)
            if x > 0:
                alpha /= x
            alpha *= pw
            shape = {"uvw": (m1, m2, mo), "uvu": (m1, m2), "uuu": (m1,)}[mode]
            if mode == "uvu":
                assert m1 == mo
            if mode == "uuu":
                assert m1 == m2 == mo
            self.instructions.append((i1, i2, io, mode, has_w, math.sqrt(alpha), shape))
        self.weight_numel = sum(math.prod(s) for *_, hw, _, s in self.instructions if hw)
        if shared_weights is False and internal_weights is None:
            internal_weights = False
        if shared_weights is None:
            shared_weights = True
        if internal_weights is None:
            internal_weights = shared_weights and self.weight_numel > 0
        self.shared_weights = shared_weights
        self.internal_weights = internal_weights
        if internal_weights and self.weight_numel > 0:
            self.weight = nn.Parameter(torch.randn(self.weight_numel))
        else:
            self.register_buffer("weight", torch.zeros(0))
        touched = {io for (_, _, io, *_r) in self.instructions}
        self.register_buffer(
            "output_mask",
            torch.cat([(torch.ones if i in touched else torch.zeros)(mul * ir.dim)
                       for i, (mul, ir) in enumerate(self.irreps_out)]) if len(self.irreps_out) else torch.zeros(0),
        )

    def forward(self, x1, x2, weight=None):
        z = x1.shape[0]
        if weight is None:
            assert self.internal_weights or self.weight_numel == 0
            weight = self.weight
        shared = weight.dim() == 1
        s1, s2 = self.irreps_in1.slices(), self.irreps_in2.slices()
        outs = [None] * len(self.irreps_out)
        off = 0
        for (i1, i2, io, mode, has_w, pw, shape) in self.instructions:
            m1, ir1 = self.irreps_in1[i1]
            m2, ir2 = self.irreps_in2[i2]
            mo, iro = self.irreps_out[io]
            a = x1[:, s1[i1]].reshape(z, m1, ir1.dim)
            b = x2[:, s2[i2]].reshape(z, m2, ir2.dim)
            C = wigner_3j(ir1.l, ir2.l, iro.l, dtype=x1.dtype, device=x1.device)
            w = None
            if has_w:
                n = math.prod(shape)
                w = weight[off : off + n].reshape(shape) if shared else weight[:, off : off + n].reshape(z, *shape)
                off += n
            xx = torch.einsum("zui,zvj->zuvij", a, b)
            if mode == "uvu":
                if w is None:
                    y = torch.einsum("ijk,zuvij->zuk", C, xx)
                elif shared:
                    y = torch.einsum("uv,ijk,zuvij->zuk", w, C, xx)
                else:
                    y = torch.einsum("zuv,ijk,zuvij->zuk", w, C, xx)
            elif mode == "uvw":
                assert w is not None
                if shared:
                    y = torch.einsum("uvw,ijk,zuvij->zwk", w, C, xx)
                else:
                    y = torch.einsum("zuvw,ijk,zuvij->zwk", w, C, xx)
            elif mode == "uuu":
                xx = torch.einsum("zui,zuj->zuij", a, b)
                if w is None:
                    y = torch.einsum("ijk,zuij->zuk", C, xx)
                elif shared:
                    y = torch.einsum("u,ijk,zuij->zuk", w, C, xx)
                else:
                    y = torch.einsum("zu,ijk,zuij->zuk", w, C, xx)
            else:
                raise NotImplementedError(mode)
            y = pw * y
            outs[io] = y if outs[io] is None else outs[io] + y
        cols = []
        for i, (mul, ir) in enumerate(self.irreps_out):
            cols.append(x1.new_zeros(z, mul * ir.dim) if outs[i] is None else outs[i].reshape(z, mul * ir.dim))
        return torch.cat(cols, dim=1) if cols else x1.new_zeros(z, 0)


class FullyConnectedTensorProduct(TensorProduct):
    """o3.FullyConnectedTensorProduct.  Reference call site:
    ``nn/message_passing.py:83-87`` (self-connection ``sc``)."""

    def __init__(self, irreps_in1, irreps_in2, irreps_out):
        irreps_in1, irreps_in2, irreps_out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        instr = [
            (i1, i2, io, "uvw", True, 1.0)
            for i1, (_, ir1) in enumerate(irreps_in1)
            for i2, (_, ir2) in enumerate(irreps_in2)
            for io, (_, iro) in enumerate(irreps_out)
            if iro in ir1 * ir2
        ]
        super().__init__(irreps_in1, irreps_in2, irreps_out, instr, shared_weights=True, internal_weights=True)


class ElementwiseTensorProduct(TensorProduct):
    def __init__(self, irreps_in1, irreps_in2):
        irreps_in1, irreps_in2 = Irreps(irreps_in1).simplify(), Irreps(irreps_in2).simplify()
        assert irreps_in1.num_irreps == irreps_in2.num_irreps
        l1, l2 = list(irreps_in1), list(irreps_in2)
        i = 0
        while i < len(l1):  # split blocks so that multiplicities line up
            (m1, ir1), (m2, ir2) = l1[i], l2[i]
            if m1 < m2:
                l2[i] = (m1, ir2)
                l2.insert(i + 1, (m2 - m1, ir2))
            if m2 < m1:
                l1[i] = (m2, ir1)
                l1.insert(i + 1, (m1 - m2, ir1))
            i += 1
        out, instr = [], []
        for i, ((mul, ir1), (mul2, ir2)) in enumerate(zip(l1, l2)):
            assert mul == mul2
            for ir in ir1 * ir2:
                instr.append((i, i, len(out), "uuu", False, 1.0))
                out.append((mul, ir))
        super().__init__(Irreps(l1), Irreps(l2), Irreps(out), instr, shared_weights=True, internal_weights=False)


# ------------------------------------------------------------------------------------------
class SphericalHarmonics(nn.Module):
    """o3.SphericalHarmonics(irreps_out, normalize, normalization).
    Reference call site: ``nn/embedding.py:163-165``."""

    def __init__(self, irreps_out, normalize, normalization="integral", irreps_in=None):
        super().__init__()
        if isinstance(irreps_out, int):
            irreps_out = Irreps.spherical_harmonics(irreps_out)
        self.irreps_out = Irreps(irreps_out)
        self._ls = []
        for mul, ir in self.irreps_out:
            assert ir.p == (-1) ** ir.l, "spherical harmonics have natural parity (1o input)"
            self._ls += [ir.l] * mul
        self.normalize = normalize
        self.normalization = normalization

    def forward(self, x):
        return spherical_harmonics(self._ls, x, self.normalize, self.normalization)


# ------------------------------------------------------------------------------------------
class _FCLayer(nn.Module):
    def __init__(self, h_in, h_out, act, var_in, var_out):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(h_in, h_out))
        self.act = act
        self.h_in, self.h_out, self.var_in, self.var_out = h_in, h_out, var_in, var_out

    def forward(self, x):
        if self.act is not None:
            w = self.weight / (self.h_in * self.var_in) ** 0.5
            x = self.act(x @ w)
            return x * self.var_out ** 0.5
        w = self.weight / (self.h_in * self.var_in / self.var_out) ** 0.5
        return x @ w


class FullyConnectedNet(nn.Sequential):
    """e3nn.nn.FullyConnectedNet: bias-free MLP, normalize2mom activations, no
    activation on the last layer.  Reference call site: ``nn/message_passing.py:74-79``."""

    def __init__(self, hs, act=None, variance_in=1, variance_out=1, out_act=False):
        super().__init__()
        self.hs = list(hs)
        if act is not None:
            act = normalize2mom(act)
        var_in = variance_in
        for i, (h1, h2) in enumerate(zip(self.hs, self.hs[1:])):
            if i == len(self.hs) - 2:
                lay = _FCLayer(h1, h2, act if out_act else None, var_in, variance_out)
            else:
                lay = _FCLayer(h1, h2, act, var_in, 1.0)
            setattr(self, f"layer{i}", lay)
            var_in = 1.0


# ------------------------------------------------------------------------------------------
class Activation(nn.Module):
    """e3nn.nn.Activation on scalar blocks."""

    def __init__(self, irreps_in, acts):
        super().__init__()
        irreps_in = Irreps(irreps_in)
        assert len(irreps_in) == len(acts)
        acts = [normalize2mom(a) if a is not None else None for a in acts]
        out = []
        for (mul, ir), act in zip(irreps_in, acts):
            if act is not None:
                assert ir.l == 0
                x = torch.linspace(0, 10, 256)
                a1, a2 = act(x), act(-x)
                if (a1 - a2).abs().max() < 1e-5:
                    p_act = 1
                elif (a1 + a2).abs().max() < 1e-5:
                    p_act = -1
                else:
                    p_act = 0
                p_out = p_act if ir.p == -1 else ir.p
                if p_out == 0:
                    raise ValueError("parity violated: odd scalar needs an even or odd activation")
                out.append((mul, (0, p_out)))
            else:
                out.append((mul, ir))
        self.irreps_in = irreps_in
        self.irreps_out = Irreps(out)
        self.acts = nn.ModuleList([a if a is not None else nn.Identity() for a in acts])

    def forward(self, x):
        cols = []
        for sl, act in zip(self.irreps_in.slices(), self.acts):
            cols.append(act(x[..., sl]))
        return torch.cat(cols, dim=-1) if cols else x


class Gate(nn.Module):
    """e3nn.nn.Gate.  Reference call site: ``nn/message_passing.py:195-205``."""

    def __init__(self, irreps_scalars, act_scalars, irreps_gates, act_gates, irreps_gated):
        super().__init__()
        irreps_scalars, irreps_gates, irreps_gated = Irreps(irreps_scalars), Irreps(irreps_gates), Irreps(irreps_gated)
        assert irreps_gates.num_irreps == irreps_gated.num_irreps
        assert all(ir.l == 0 for _, ir in irreps_gates) and all(ir.l == 0 for _, ir in irreps_scalars)
        self.irreps_scalars, self.irreps_gates, self.irreps_gated = irreps_scalars, irreps_gates, irreps_gated
        self._irreps_in = (irreps_scalars + irreps_gates + irreps_gated).simplify()
        self.act_scalars = Activation(irreps_scalars, act_scalars)
        self.act_gates = Activation(irreps_gates, act_gates)
        self.mul = ElementwiseTensorProduct(irreps_gated, self.act_gates.irreps_out)
        self._irreps_out = self.act_scalars.irreps_out + self.mul.irreps_out

    @property
    def irreps_in(self):
        return self._irreps_in

    @property
    def irreps_out(self):
        return self._irreps_out

    def forward(self, features):
        ns, ng = self.irreps_scalars.dim, self.irreps_gates.dim
        scalars = features[..., :ns]
        gates = features[..., ns : ns + ng]
        gated = features[..., ns + ng :]
        scalars = self.act_scalars(scalars)
        if gates.shape[-1]:
            gates = self.act_gates(gates)
            gated = self.mul(gated, gates)
            return torch.cat([scalars, gated], dim=-1)
        return scalars


class NormActivation(nn.Module):
    """Never selected by an in-scope config; constructing it is an error here."""

    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("NormActivation is outside the hot path (SURVEY A.7)")


# ------------------------------------------------------------------------------------------
def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    """torch_runstats.scatter.scatter 0.2.0: zeros().scatter_add_(), sum only.
    Reference call sites: ``nn/message_passing.py:109``, ``nn/output.py:69``."""
    assert reduce == "sum"
    assert dim == 0
    index = index.reshape(-1)
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[0] = dim_size
    res = src.new_zeros(shape)
    idx = index.reshape(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return res.scatter_add_(0, idx, src)


def soft_one_hot_linspace(*a, **k):  # imported (unused) by nn/embedding.py:21
    raise NotImplementedError
