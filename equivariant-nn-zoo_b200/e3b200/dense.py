"""Dense contractions of the interaction block: per-irrep linear maps (e3nn ``o3.Linear``), the
bias-free radial MLP (e3nn ``nn.FullyConnectedNet``) and the scalar-attribute self-connection
(e3nn ``o3.FullyConnectedTensorProduct`` with an all-0e second operand).

Parameter names, shapes and flat layouts follow e3nn 0.4.4 so that reference checkpoints load
(SURVEY.md A.9).  These are real GEMMs; this module issues them as library GEMMs (cuBLAS via
torch.matmul, fp32 'highest' precision) while accepting / producing the channel-fastest "imu"
layout of the fused kernels directly, so no separate transpose pass is needed.  The tcgen05
3xTF32 replacements plug in behind the same classes.
"""
import math

import torch
from torch import nn

from . import ops
from .irreps import Irreps

# normalize2mom constants (e3nn estimates E[act(z)^2]^-1/2 by Monte-Carlo with 1e6 float64
# samples from torch.Generator('cpu').manual_seed(0); values of that recipe)
ACT_CST = {
    "ssp": 1.878204668541552,
    "silu": 1.6791767923989418,
    "tanh": 1.5937334472592692,
    "tanhlu": 1.1501944455736173,
    "abs": 1.001110600838467,
}
# parity of each activation as a function: +1 even, -1 odd, 0 neither
ACT_PARITY = {"ssp": 0, "silu": 0, "tanh": -1, "tanhlu": -1, "abs": 1}


def _mm(x, W, alpha):
    """alpha * x @ W.  In second-order mode the product is a tcgen05 node whose backward is differentiable again
    (ops.dense); otherwise a library GEMM."""
    if ops.second_order_active() and ops.dense_supported(x, W):
        return ops.dense(x, W, alpha)
    return torch.matmul(x, W * alpha)


def _mask(irreps, touched):
    parts = [(torch.ones if i in touched else torch.zeros)(b.dim) for i, b in enumerate(irreps)]
    return torch.cat(parts) if parts else torch.zeros(0)


class Linear(nn.Module):
    """y[z, w, m] = alpha * sum_u W[u, w] x[z, u, m] between equal irreps; alpha = 1/sqrt(fan_in)
    summed over all input blocks feeding an output block; optional bias on 0e outputs.
    ``in_layout`` / ``out_layout``: "mul_ir" (e3nn) or "imu" (channel fastest)."""

    def __init__(self, irreps_in, irreps_out, biases=False, in_layout="mul_ir", out_layout="mul_ir"):
        super().__init__()
        self.irreps_in, self.irreps_out = Irreps(irreps_in), Irreps(irreps_out)
        self.in_layout, self.out_layout = in_layout, out_layout
        pairs = [(i, o) for i, a in enumerate(self.irreps_in) for o, b in enumerate(self.irreps_out) if a.ir == b.ir]
        fan = {}
        for i, o in pairs:
            fan[o] = fan.get(o, 0) + self.irreps_in[i].mul
        self.paths = []
        off = 0
        for i, o in pairs:
            mi, mo = self.irreps_in[i].mul, self.irreps_out[o].mul
            self.paths.append((i, o, off, 1.0 / math.sqrt(fan[o]) if fan[o] > 0 else 1.0))
            off += mi * mo
        self.weight_numel = off
        self.weight = nn.Parameter(torch.randn(off))
        self.bias_blocks = [o for o, b in enumerate(self.irreps_out) if biases and b.ir.is_scalar()]
        nb = sum(self.irreps_out[o].mul for o in self.bias_blocks)
        if nb:
            self.bias = nn.Parameter(torch.zeros(nb))
        else:
            self.register_buffer("bias", torch.zeros(0))
        self.register_buffer("output_mask", _mask(self.irreps_out, {o for _, o, _, _ in self.paths} | set(self.bias_blocks)))
        self._in_slices, self._out_slices = self.irreps_in.slices(), self.irreps_out.slices()
        self._spec = ops.SCSpec(self.irreps_in, self.irreps_out, 0, list(self.paths))

    def _block_in(self, x, i):
        """-> [z, d, mul] view/copy of input block i (contraction index last)"""
        b = self.irreps_in[i]
        blk = x[:, self._in_slices[i]]
        if self.in_layout == "imu":
            return blk.reshape(-1, b.ir.dim, b.mul)
        return blk.reshape(-1, b.mul, b.ir.dim).transpose(1, 2)

    def _forward_node(self, x):
        """second-order mode: the whole map as ONE bilinear tcgen05 node (grouped launch over the irreps blocks)"""
        xi = x if self.in_layout == "imu" else ops.layout(x, self.irreps_in, True)
        y = ops.block_linear(xi, self.weight, self._spec)
        if self.bias_blocks:
            cols, boff = [], 0
            for o, blk in enumerate(self.irreps_out):
                if o in self.bias_blocks:
                    cols.append(self.bias[boff:boff + blk.mul])
                    boff += blk.mul
                else:
                    cols.append(y.new_zeros(blk.dim))
            y = y + torch.cat(cols)                      # scalar blocks: imu and mul_ir coincide
        return y if self.out_layout == "imu" else ops.layout(y, self.irreps_out, False)

    def forward(self, x):
        z = x.shape[0]
        if (ops.second_order_active() and self.paths and self._spec.ok and z > 0
                and ((x.is_cuda and x.dtype == torch.float32) or ops.FORCE_DENSE_FUNCTION)):
            return self._forward_node(x)
        if x.is_cuda and x.requires_grad and self.in_layout == "mul_ir" and not ops.second_order_active():
            live = sorted({i for i, _, _, _ in self.paths})
            if 0 < len(live) < len(self.irreps_in):
                x = ops.tag_live_blocks(x, live, self.irreps_in)      # the gradient of the other blocks is exactly zero
        acc = [None] * len(self.irreps_out)
        for i, o, off, alpha in self.paths:
            mi, mo = self.irreps_in[i].mul, self.irreps_out[o].mul
            W = self.weight[off:off + mi * mo].reshape(mi, mo)
            y = _mm(self._block_in(x, i), W, alpha)        # [z, d, mo]
            acc[o] = y if acc[o] is None else acc[o] + y
        boff = 0
        for o in self.bias_blocks:
            mo = self.irreps_out[o].mul
            b = self.bias[boff:boff + mo].reshape(1, 1, mo)
            boff += mo
            acc[o] = b.expand(z, 1, mo) if acc[o] is None else acc[o] + b
        cols = []
        for o, blk in enumerate(self.irreps_out):
            if acc[o] is None:
                cols.append(x.new_zeros(z, blk.dim))
            elif self.out_layout == "imu":
                cols.append(acc[o].reshape(z, blk.dim))
            else:
                cols.append(acc[o].transpose(1, 2).reshape(z, blk.dim))
        return torch.cat(cols, dim=1) if len(cols) != 1 else cols[0]


class _DenseLayer(nn.Module):
    def __init__(self, h_in, h_out):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(h_in, h_out))


def ssp(x):
    return torch.nn.functional.softplus(x) - math.log(2.0)


class RadialMLP(nn.Module):
    """e3nn nn.FullyConnectedNet(hs, ssp): h <- c_ssp * ssp(h @ W / sqrt(fan_in)) on hidden
    layers, plain h @ W / sqrt(fan_in) on the last; no biases.  Parameters layer{i}.weight."""

    def __init__(self, hs, act="ssp"):
        super().__init__()
        self.hs = list(hs)
        assert act == "ssp"
        self.cst = ACT_CST[act]
        for i, (a, b) in enumerate(zip(self.hs, self.hs[1:])):
            setattr(self, f"layer{i}", _DenseLayer(a, b))
        self.n_layers = len(self.hs) - 1

    def forward(self, h):
        for i in range(self.n_layers):
            W = getattr(self, f"layer{i}").weight
            h = _mm(h, W, 1.0 / math.sqrt(W.shape[0]))
            if i < self.n_layers - 1:
                h = ssp(h) * self.cst
        return h


class ScalarAttrTensorProduct(nn.Module):
    """FullyConnectedTensorProduct(features, node_attrs (all 0e), out): with scalar attributes
    the CG tensor is delta/sqrt(2l+1) and cancels the component normalisation, leaving
    out[z, w, m] = alpha * sum_{u, v} W[u, v, w] x[z, u, m] a[z, v],  alpha = 1/sqrt(sum mul1*mul2)
    over the instructions feeding an output block.  Flat weight: instructions in (i1, i2, i_out)
    loop order, each [mul1, mul2, mul_out] (e3nn layout)."""

    def __init__(self, irreps_in1, irreps_in2, irreps_out):
        super().__init__()
        self.irreps_in1, self.irreps_in2, self.irreps_out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        if any(not b.ir.is_scalar() for b in self.irreps_in2):
            raise NotImplementedError("self-connection attributes must be 0e scalars (all reference configs)")
        ins = [(i1, i2, o) for i1, a in enumerate(self.irreps_in1) for i2, _ in enumerate(self.irreps_in2)
               for o, c in enumerate(self.irreps_out) if c.ir == a.ir]
        fan = {}
        for i1, i2, o in ins:
            fan[o] = fan.get(o, 0) + self.irreps_in1[i1].mul * self.irreps_in2[i2].mul
        self.paths, off = [], 0
        for i1, i2, o in ins:
            n = self.irreps_in1[i1].mul * self.irreps_in2[i2].mul * self.irreps_out[o].mul
            self.paths.append((i1, i2, o, off, 1.0 / math.sqrt(fan[o])))
            off += n
        self.weight_numel = off
        self.weight = nn.Parameter(torch.randn(off))
        self.register_buffer("output_mask", _mask(self.irreps_out, {o for _, _, o, _, _ in self.paths}))
        self._s1, self._s2 = self.irreps_in1.slices(), self.irreps_in2.slices()
        self._spec = None
        if len(self.irreps_in2) == 1:
            self._spec = ops.SCSpec(self.irreps_in1, self.irreps_out, self.irreps_in2.dim,
                                    [(i1, o, off, alpha) for i1, _, o, off, alpha in self.paths])

    def forward(self, x, attrs):
        """x mul_ir [z, in1.dim], attrs [z, in2.dim] -> mul_ir [z, out.dim]"""
        z = x.shape[0]
        if (ops.second_order_active() and self._spec is not None and self._spec.ok and z > 0
                and ((x.is_cuda and x.dtype == torch.float32) or ops.FORCE_DENSE_FUNCTION)):
            # second-order mode: trilinear tcgen05 nodes on the channel-fastest layout, no (u, v) outer product
            y = ops.self_connection(ops.layout(x, self.irreps_in1, True), attrs, self.weight, self._spec)
            return ops.layout(y, self.irreps_out, False)
        acc = [None] * len(self.irreps_out)
        for i1, i2, o, off, alpha in self.paths:
            m1, m2, mo = self.irreps_in1[i1].mul, self.irreps_in2[i2].mul, self.irreps_out[o].mul
            d = self.irreps_in1[i1].ir.dim
            W = self.weight[off:off + m1 * m2 * mo].reshape(m1 * m2, mo)
            xb = x[:, self._s1[i1]].reshape(z, m1, d)
            ab = attrs[:, self._s2[i2]]
            # outer product over (u, v) -> K = m1*m2, then one GEMM per instruction
            xa = (xb.unsqueeze(2) * ab.reshape(z, 1, m2, 1)).reshape(z, m1 * m2, d).transpose(1, 2)  # [z, d, K]
            y = _mm(xa, W, alpha)  # [z, d, mo]
            acc[o] = y if acc[o] is None else acc[o] + y
        cols = []
        for o, blk in enumerate(self.irreps_out):
            cols.append(x.new_zeros(z, blk.dim) if acc[o] is None else acc[o].transpose(1, 2).reshape(z, blk.dim))
        return torch.cat(cols, dim=1)
