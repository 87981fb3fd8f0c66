"""Node/edge-wise layers (reference ``e3_layers/nn/pointwise.py``)."""
import torch

from e3b200 import dense, ops
from e3b200.irreps import Irreps
from e3b200.plan import TPStructure

from .sequential import Module


class PointwiseLinear(Module):
    def __init__(self, irreps_in, irreps_out, biases=True, **kwargs):
        super().__init__()
        self.init_irreps(input=irreps_in, output=irreps_out, output_keys=["output"])
        self.linear = dense.Linear(self.irreps_in["input"], self.irreps_out["output"], biases=biases)

    def forward(self, data, attrs):
        x = data["input"]
        out = self.linear(x)
        tag = getattr(x, "_e3b_onehot", None)
        if tag is not None:
            # rows of `out` are rows of the table linear(identity): an interaction block that takes `out` as its node
            # attributes contracts its self-connection weights with the table once per species instead of once per node
            out._e3b_species = (tag[0], tag[1], self.linear)
        return ({"output": out},
                {"output": (attrs["input"][0], self.irreps_out["output"])})


class LayerNormalization(Module):
    """per irrep block: x / sqrt(sum x^2 / mul + 1e-6) * std_block (reference nn/pointwise.py:32-51).
    One kernel forward, one backward; in second-order mode (graph of the gradient) the closed form below
    is differentiated by torch instead."""

    def __init__(self, irreps_in, irreps_out, **kwargs):
        super().__init__()
        self.init_irreps(input=irreps_in, output=irreps_out, output_keys=["output"])
        assert irreps_in == irreps_out
        irr = Irreps(irreps_in)
        self.muls = [b.mul for b in irr]
        self.ls = [b.ir.l for b in irr]
        self.slices = [(s.start, s.stop) for s in irr.slices()]
        self.std = torch.nn.Parameter(torch.ones(len(self.slices)))

    def forward(self, data, attrs):
        x = data["input"]
        if x.is_cuda and not ops.second_order_active():
            return {"output": ops.layer_norm(x, self.std, self.muls, self.ls, 1e-6)}, attrs
        cols = []
        for b, (lo, hi) in enumerate(self.slices):
            blk = x[:, lo:hi]
            rms = ((blk * blk).sum(dim=-1, keepdim=True) / self.muls[b] + 1e-6) ** 0.5
            cols.append(blk / rms * self.std[b])
        return {"output": torch.cat(cols, dim=1)}, attrs


class _PathTable(torch.nn.Module):
    """stands where the reference keeps its e3nn TensorProduct (``self.tp.tp``): exposes
    ``weight_numel`` / ``irreps_out`` and the empty ``weight`` / ``output_mask`` buffers of an
    externally weighted product so state_dict keys line up."""

    def __init__(self, structure):
        super().__init__()
        self.structure = structure
        self.weight_numel = structure.weight_numel
        self.irreps_in1, self.irreps_in2, self.irreps_out = structure.irreps_in, structure.irreps_sh, structure.irreps_mid
        self.register_buffer("weight", torch.zeros(0))
        self.register_buffer("output_mask", torch.ones(structure.irreps_mid.dim))


class TensorProductExpansion(Module):
    """uvu tensor product of node features with edge spherical harmonics, externally weighted,
    followed by the per-irrep linear map.  On the B200 path the product, the gather by source and
    the sum over incoming edges are ONE kernel (``ops.tp_conv``) and the linear map is applied
    after the reduction (it commutes with the edge sum)."""

    def __init__(self, left, right, output, instruction="uvu", internal_weight=True, **kwargs):
        super().__init__()
        self.init_irreps(left=left, right=right, output=output, output_keys=["output"])
        if instruction != "uvu" or internal_weight:
            raise NotImplementedError("B200 TensorProductExpansion implements the externally weighted 'uvu' case "
                                      "(the only one on the interaction-block path)")
        st = TPStructure(self.irreps_in["left"], self.irreps_in["right"], self.irreps_out["output"])
        self.tp = _PathTable(st)
        self.internal_weight = internal_weight
        self.linear = dense.Linear(st.irreps_mid.simplify(), self.irreps_out["output"], in_layout="imu")
        self._plan = None

    @property
    def plan(self):
        if self._plan is None:
            self._plan = ops.TPPlan(self.tp.structure)
        return self._plan

    def forward(self, left=None, right=None, weight=None, csr=None):
        """left: node features in imu layout [N, .]; right [E, sh]; weight [E, weight_numel]"""
        mid = ops.tp_conv(left, right, weight, self.plan, csr)
        return self.linear(mid)


class Concat(Module):
    def __init__(self, irreps_out, **irreps_in):
        super().__init__()
        self.init_irreps(**irreps_in, output=irreps_out, output_keys=["output"])
        total = Irreps([])
        for v in self.irreps_in.values():
            total = total + Irreps(v)
        self.linear = dense.Linear(total, Irreps(self.irreps_out["output"]), biases=True)

    def forward(self, data, attrs):
        keys = list(self.irreps_in.keys())
        x = torch.cat([data[k] for k in keys], dim=1)
        return {"output": self.linear(x)}, {"output": (attrs[keys[0]][0], self.irreps_out["output"])}
