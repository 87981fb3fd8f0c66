"""The Nequip-style interaction block (reference ``e3_layers/nn/message_passing.py``) on the
B200 path.

``FactorizedConvolution``: radial MLP -> per-edge weights; ``linear_1``; then ONE fused kernel for
gather-by-source + weighted uvu Clebsch-Gordan product + sum over the destination's incoming edges
(CSR segments, no atomics); the per-irrep linear map runs after the reduction; ``/ sqrt(avg_num_neighbors)``;
self-connection.  ``MessagePassing`` adds the gate non-linearity (kernel), optional resnet and
LayerNormalization.  Constructor kwargs, forward protocol and parameter names follow the reference."""
import ctypes
import math
import os

import torch

from e3b200 import _lib, dense, interaction, ops
from e3b200.irreps import Irreps

from ..utils import activation_name, build, tp_path_exists
from .pointwise import LayerNormalization, TensorProductExpansion
from .sequential import Module


# E3B_FUSED=0 composes the block op by op (A/B comparison of the two fp32 paths; both are CUDA kernels)
FUSED_BLOCKS = os.environ.get("E3B_FUSED", "1") != "0"


class FactorizedConvolution(Module):
    def __init__(self, input_features, output_features, node_attrs, edge_radial, edge_spherical,
                 invariant_layers=1, invariant_neurons=8, avg_num_neighbors=None, use_sc=True,
                 nonlinearity_scalars={"e": "ssp"}, reduce=True):
        super().__init__()
        self.init_irreps(input_features=input_features, output_features=output_features, node_attrs=node_attrs,
                         edge_radial=edge_radial, edge_spherical=edge_spherical, output_keys=["output_features"])
        if not reduce:
            raise NotImplementedError("reduce=False (per-edge output) is only used by the out-of-scope Pairwise head")
        self.avg_num_neighbors, self.use_sc, self.reduce = avg_num_neighbors, use_sc, reduce
        feat_in = self.irreps_in["input_features"]
        feat_out = self.irreps_out["output_features"]
        # linear_1 writes the channel-fastest layout the fused kernel gathers from
        self.linear_1 = dense.Linear(feat_in, feat_in, out_layout="imu")
        self.tp = TensorProductExpansion(feat_in, (self.irreps_in["edge_spherical"], "edge_spherical"),
                                         (feat_out, "edge_features"), "uvu", internal_weight=False)
        n_radial = Irreps(self.irreps_in["edge_radial"]).num_irreps
        self.fc = dense.RadialMLP([n_radial] + invariant_layers * [invariant_neurons] + [self.tp.tp.weight_numel], "ssp")
        self.sc = dense.ScalarAttrTensorProduct(feat_in, Irreps(self.irreps_in["node_attrs"]), feat_out) if use_sc else None

    def forward(self, data, attrs):
        x = data["input_features"]
        edge_index = data["edge_index"]
        csr = ops.graph_of(edge_index, x.shape[0])
        weight = self.fc(data["edge_radial"])
        sc = self.sc(x, data["node_attrs"]) if self.sc is not None else None
        out = self.tp(left=self.linear_1(x), right=data["edge_spherical"], weight=weight, csr=csr)
        if self.avg_num_neighbors is not None:
            out = out * (1.0 / math.sqrt(self.avg_num_neighbors))
        if sc is not None:
            out = out + sc
        return ({"output_features": out},
                {"output_features": (attrs["input_features"][0], self.irreps_out["output_features"])})


class _Gate(torch.nn.Module):
    """e3nn nn.Gate as one kernel: input [scalars | gates | gated] -> [act(scalars) | gated * act(gates)]"""

    def __init__(self, irreps_scalars, act_scalars, irreps_gates, act_gates, irreps_gated):
        super().__init__()
        self.irreps_scalars, self.irreps_gates, self.irreps_gated = irreps_scalars, irreps_gates, irreps_gated
        assert len(irreps_gates) == len(irreps_gated) and all(g.mul == b.mul for g, b in zip(irreps_gates, irreps_gated))
        d = _lib.GateDesc()
        out_scalars = []
        d.n_scalar_blocks = len(irreps_scalars)
        for i, (blk, act) in enumerate(zip(irreps_scalars, act_scalars)):
            par = dense.ACT_PARITY[act]
            p_out = blk.ir.p if blk.ir.p == 1 else par
            if p_out == 0:
                raise ValueError(f"activation {act} has no definite parity: cannot act on {blk}")
            out_scalars.append((blk.mul, (0, p_out)))
            d.scalar_mul[i], d.scalar_act[i], d.scalar_cst[i] = blk.mul, ops.ACT_CODES[act], self._cst(act)
        d.n_gated_blocks = len(irreps_gated)
        for i, (blk, act) in enumerate(zip(irreps_gated, act_gates)):
            assert irreps_gates[i].ir.is_scalar(), "gates must be 0e scalars"
            d.gated_mul[i], d.gated_l[i], d.gate_act[i], d.gate_cst[i] = blk.mul, blk.ir.l, ops.ACT_CODES[act], self._cst(act)
        self.desc = d
        self.irreps_in = (irreps_scalars + irreps_gates + irreps_gated).simplify()
        self.irreps_out = Irreps(out_scalars) + irreps_gated

    @staticmethod
    def _cst(act):
        c = dense.ACT_CST[act]
        return 1.0 if abs(c - 1.0) < 1e-4 else c

    def forward(self, x):
        return ops.gate(x, self.desc, self.irreps_out.dim)


class MessagePassing(Module):
    def __init__(self, input_features, output_features, node_attrs, edge_radial, edge_spherical, convolution,
                 resnet=False, nonlinearity_type="gate", nonlinearity_scalars={"e": "ssp", "o": "tanh"},
                 nonlinearity_gates={"e": "ssp", "o": "abs"}, normalize=False):
        super().__init__()
        self.init_irreps(input_features=input_features, output_features=output_features, node_attrs=node_attrs,
                         edge_radial=edge_radial, edge_spherical=edge_spherical, output_keys=["output_features"])
        if nonlinearity_type != "gate":
            raise NotImplementedError("only the gate non-linearity is on the accelerated path (no config uses 'norm')")
        by_parity_s = {1: activation_name(nonlinearity_scalars["e"]), -1: activation_name(nonlinearity_scalars["o"])}
        by_parity_g = {1: activation_name(nonlinearity_gates["e"]), -1: activation_name(nonlinearity_gates["o"])}
        sh = Irreps(self.irreps_in["edge_spherical"])
        prev = Irreps(self.irreps_in["input_features"])
        hidden = Irreps(self.irreps_out["output_features"])
        self.feature_irreps_hidden = hidden
        reach = [b for b in hidden if tp_path_exists(prev, sh, b.ir)]
        scalars = Irreps([b for b in reach if b.ir.l == 0])
        gated = Irreps([b for b in reach if b.ir.l > 0])
        gates = Irreps([(b.mul, "0e") for b in gated])
        self.equivariant_nonlin = _Gate(scalars, [by_parity_s[b.ir.p] for b in scalars],
                                        gates, [by_parity_g[b.ir.p] for b in gates], gated)
        conv_out = self.equivariant_nonlin.irreps_in.simplify()
        self.resnet = bool(resnet and (scalars + gated).simplify() == prev)
        self.conv = build(convolution, input_features=input_features, output_features=conv_out, node_attrs=node_attrs,
                          edge_radial=edge_radial, edge_spherical=edge_spherical)
        self._fused = None
        self.normalize = normalize
        if normalize:
            self.norm = LayerNormalization(self.irreps_out["output_features"], self.irreps_out["output_features"])

    @property
    def fused(self):
        """description of this block for the single-node fp32 path (e3b200.interaction)"""
        if self._fused is None:
            self._fused = interaction.FusedInteraction(self)
        return self._fused

    def forward(self, data, attrs):
        skip = data["input_features"]
        if (skip.dtype == torch.float32 and skip.is_cuda and self.fused.reason is None and FUSED_BLOCKS
                and not ops.second_order_active()):
            edge_index = data["edge_index"]
            csr = ops.graph_of(edge_index, skip.shape[0])
            out, out_imu = interaction.interaction(self.fused, skip, data["node_attrs"], data["edge_radial"],
                                                   data["edge_spherical"], csr, group=getattr(self, "_e3b_group", None),
                                                   edge_index=edge_index)
            out._e3b_imu = out_imu       # the next block gathers from the channel-fastest twin
        else:
            # fp64 correctness mode / irregular irreps / second-order mode (graph of the gradient): the same
            # kernels composed op by op
            out = self.conv(data, attrs)[0]["output_features"]
            out = self.equivariant_nonlin(out)
        if self.resnet:
            out = skip + out
        if self.normalize:
            out = self.norm({"input": out}, attrs)[0]["output"]
        return ({"output_features": out},
                {"output_features": (attrs["input_features"][0], self.irreps_out["output_features"])})
