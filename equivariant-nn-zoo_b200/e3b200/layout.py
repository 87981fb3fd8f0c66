"""mul_ir (e3nn: [block][u][m]) <-> imu (library: [block][m][u]) conversions in plain torch.
Host-side helper for weights/tests; activations are converted inside the dense contractions."""
import torch

from .irreps import Irreps


def to_imu(x, irreps):
    irreps = Irreps(irreps)
    cols = []
    for sl, (mul, ir) in zip(irreps.slices(), irreps):
        cols.append(x[..., sl].reshape(*x.shape[:-1], mul, ir.dim).transpose(-1, -2).reshape(*x.shape[:-1], mul * ir.dim))
    return torch.cat(cols, dim=-1) if cols else x


def from_imu(x, irreps):
    irreps = Irreps(irreps)
    cols = []
    for sl, (mul, ir) in zip(irreps.slices(), irreps):
        cols.append(x[..., sl].reshape(*x.shape[:-1], ir.dim, mul).transpose(-1, -2).reshape(*x.shape[:-1], mul * ir.dim))
    return torch.cat(cols, dim=-1) if cols else x
