"""Output heads (reference ``e3_layers/nn/output.py:18-74``): gradients of a scalar output with
respect to positions (forces / scores) and pooling of node features into graph features."""
import torch

from e3b200 import ops
from e3b200.irreps import Irreps

from ..utils import ConfigDict, build
from .sequential import Module, strip_reference_state


class GradientOutput(Module):
    """gradients = sign * d(sum y)/dx through the wrapped network (hand-written backward kernels).
    In training mode the reference builds the graph of the gradient as well (``create_graph=
    self.training``, nn/output.py:39-43) so that force-matching losses can be differentiated with
    respect to the parameters: the network then runs in second-order mode (``ops.second_order``),
    where every backward kernel is itself a differentiable node."""

    def __init__(self, func, x, y, gradients, sign: float = 1.0, **kwargs):
        super().__init__()
        self.sign = float(sign)
        assert self.sign in (1.0, -1.0)
        self.init_irreps(x=x, y=y, gradients=gradients, output_keys=["gradients"])
        assert Irreps(self.irreps_in["y"]).lmax == 0
        self.func = build(func, **kwargs) if isinstance(func, (dict, ConfigDict)) else func

    def load_state_dict(self, state_dict, strict=True, **kwargs):
        """accepts the reference's checkpoints as they are (``sequential.strip_reference_state``)"""
        return super().load_state_dict(strip_reference_state(state_dict, set(self.state_dict().keys())), strict=strict, **kwargs)

    def forward(self, data):
        create_graph = bool(self.training and torch.is_grad_enabled())
        wrt = self.inputKeyMap(data)["x"]
        was = wrt.requires_grad
        wrt.requires_grad_(True)
        with torch.enable_grad():
            if create_graph:
                with ops.second_order():
                    out = self.func(data)
            else:
                out = self.func(data)
            y = self.inputKeyMap(out)["y"]
            with ops.positions_only(wrt):                    # this backward pass needs no parameter gradients
                (grad,) = torch.autograd.grad(y.sum(), wrt, create_graph=create_graph)
        wrt.requires_grad_(was)
        is_per = self.inputKeyMap(data.attrs)["x"][0]
        out.attrs.update(self.outputKeyMap({"gradients": (is_per, self.irreps_out["gradients"])}))
        out.update(self.outputKeyMap({"gradients": self.sign * grad}))
        return out


class Pooling(Module):
    """node features -> graph features (segmented sum over each graph's consecutive nodes)"""

    def __init__(self, irreps_in, irreps_out, reduce):
        super().__init__()
        self.init_irreps(input=irreps_in, output=irreps_out, output_keys=["output"])
        assert reduce in ("sum", "mean")
        if reduce != "sum":
            raise NotImplementedError("the reference's scatter only implements 'sum' (torch_runstats 0.2.0)")
        self.reduce = reduce

    def forward(self, data, attrs):
        x = data["input"]
        counts = data["_n_nodes"].reshape(-1).to(x.device)
        seg_ptr = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=x.device)
        torch.cumsum(counts, 0, out=seg_ptr[1:])
        out = ops.segment_sum(x, seg_ptr, data["_node_segment"].to(x.device), counts.numel())
        return {"output": out}, {"output": ("graph", self.irreps_out["output"])}
