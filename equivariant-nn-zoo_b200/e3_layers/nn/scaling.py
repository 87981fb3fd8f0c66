"""Per-species scale / shift of per-atom outputs (reference ``e3_layers/nn/scaling.py``)."""
from typing import List, Optional

import torch

from .sequential import Module


def _per_type(value, num_types):
    t = torch.as_tensor(value, dtype=torch.get_default_dtype()).reshape(-1)
    if t.numel() == 1:
        t = t.expand(num_types).clone()
    assert t.shape == (num_types,), f"expected {num_types} per-type values, got {tuple(t.shape)}"
    return t


class PerTypeScaleShift(Module):
    def __init__(self, num_types: int, shifts: Optional[List[float]], scales: Optional[List[float]],
                 scales_trainable: bool = False, shifts_trainable: bool = False,
                 irreps_in="1x0e", irreps_out="1x0e", species="1x0e"):
        super().__init__()
        self.num_types = num_types
        self.init_irreps(input=irreps_in, output=irreps_out, species=species, output_keys=["output"])
        self.has_shifts, self.has_scales = shifts is not None, scales is not None
        for name, value, trainable in (("shifts", shifts, shifts_trainable), ("scales", scales, scales_trainable)):
            if value is None:
                continue
            t = _per_type(value, num_types)
            if trainable:
                setattr(self, name, torch.nn.Parameter(t))
            else:
                self.register_buffer(name, t)
        self.shifts_trainable, self.scales_trainable = shifts_trainable, scales_trainable

    def forward(self, data, attrs):
        species, x = data["species"].reshape(-1), data["input"]
        if self.has_scales:
            x = self.scales.to(x.dtype)[species].view(-1, 1) * x
        if self.has_shifts:
            x = self.shifts.to(x.dtype)[species].view(-1, 1) + x
        return {"output": x}, {"output": (attrs["input"][0], self.irreps_out["output"])}
