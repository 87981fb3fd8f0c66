"""Embeddings feeding the interaction blocks (reference ``e3_layers/nn/embedding.py``): edge
spherical harmonics and the Bessel x cutoff radial basis run as libe3b200 kernels."""
import math

import torch
from torch import nn

from e3b200 import ops
from e3b200.irreps import Irreps

from ..utils import build
from .sequential import Module


def symmetricCutoff(x, factor, p=6.0):
    x = x * factor
    return (x - 1) ** 2 * (x + 1) ** 2 * (abs(x) < 1.0).float()


def _poly_cutoff(x, factor, p=6.0):
    x = x * factor
    out = 1.0 - ((p + 1.0) * (p + 2.0) / 2.0) * torch.pow(x, p) + p * (p + 2.0) * torch.pow(x, p + 1.0) \
        - (p * (p + 1.0) / 2) * torch.pow(x, p + 2.0)
    return out * (x < 1.0)


class PolynomialCutoff(nn.Module):
    def __init__(self, r_max, p=6, cutoff=_poly_cutoff):
        super().__init__()
        assert p >= 2.0
        self.p, self._factor, self.cutoff = float(p), 1.0 / float(r_max), cutoff

    def forward(self, x):
        return self.cutoff(x, self._factor, p=self.p)


class BesselBasis(nn.Module):
    """holds the (trainable) frequencies; evaluation is fused with the cutoff in the kernel"""

    def __init__(self, r_max, r_min=0, num_basis=8, trainable=True, one_over_r=True):
        super().__init__()
        self.trainable, self.num_basis, self.one_over_r = trainable, num_basis, one_over_r
        self.r_max, self.r_min = float(r_max), float(r_min)
        self.prefactor = 2.0 / (self.r_max - self.r_min)
        freqs = torch.linspace(start=1.0, end=num_basis, steps=num_basis) * math.pi
        if trainable:
            self.bessel_weights = nn.Parameter(freqs)
        else:
            self.register_buffer("bessel_weights", freqs)


class SphericalEncoding(Module):
    """vectors [cat, mul*3] -> real spherical harmonics, 'component' normalised, l <= 2."""

    def __init__(self, irreps_out, edge_sh_normalization="component", edge_sh_normalize=True, irreps_in="1x1o"):
        super().__init__()
        self.init_irreps(vectors=irreps_in, spherical_harmonics=irreps_out, output_keys=["spherical_harmonics"])
        self.mul = Irreps(self.irreps_in["vectors"])[0].mul
        out = Irreps(self.irreps_out["spherical_harmonics"])
        ls = [b.ir.l for b in out]
        if edge_sh_normalization != "component" or ls != list(range(len(ls))) or \
                any(b.mul != self.mul or b.ir.p != (-1) ** b.ir.l for b in out):
            raise NotImplementedError("B200 SphericalEncoding covers l = 0..lmax (natural parity), 'component'")
        self.lmax, self.normalize = len(ls) - 1, edge_sh_normalize

    def forward(self, data, attrs):
        v = data["vectors"]
        cat = v.shape[0]
        sh = ops.spherical_harmonics(v.reshape(cat * self.mul, 3), self.lmax, self.normalize)
        if self.mul > 1:  # [cat, mul, (lmax+1)^2] -> blocks of mul x l
            sh = sh.view(cat, self.mul, -1)
            sh = torch.cat([sh[:, :, l * l:(l + 1) ** 2].reshape(cat, -1) for l in range(self.lmax + 1)], dim=1)
        return ({"spherical_harmonics": sh},
                {"spherical_harmonics": ("edge", self.irreps_out["spherical_harmonics"])})


class RadialBasisEncoding(Module):
    """x -> Bessel(x) * cutoff(x); also used for time and relative-position embeddings."""

    def __init__(self, r_max, trainable, irreps_out, r_min=0, polynomial_degree=6, basis=BesselBasis,
                 cutoff=_poly_cutoff, irreps_in="1x0e", one_over_r=True):
        super().__init__()
        self.init_irreps(input=irreps_in, radial_embedding=irreps_out, output_keys=["radial_embedding"])
        num_basis = Irreps(self.irreps_out["radial_embedding"])[0].mul
        if basis is not BesselBasis:
            raise NotImplementedError("only the Bessel basis is implemented on the B200 path")
        name = getattr(cutoff, "__name__", str(cutoff))
        if name not in ("_poly_cutoff", "symmetricCutoff"):
            raise NotImplementedError(f"cutoff {name} has no kernel")
        self.cutoff_kind = 1 if name == "symmetricCutoff" else 0
        self.basis = basis(r_max, r_min, num_basis, trainable, one_over_r=one_over_r)
        self.cutoff = PolynomialCutoff(r_max, p=polynomial_degree, cutoff=cutoff)
        self.r_max = r_max

    def forward(self, data, attrs):
        x = data["input"]
        b = self.basis
        emb = ops.radial_basis(x.reshape(-1), b.bessel_weights, b.r_max, b.r_min, b.one_over_r, self.cutoff_kind,
                               self.cutoff.p)
        emb = emb.view(x.shape[0], -1)
        if getattr(x, "_e3b_edge_length", False):
            # a function of the edge length only: the interaction blocks may evaluate their radial MLPs once per
            # undirected edge of a symmetric graph (e3b200.interaction.undirected); any other provenance does not get the tag
            emb._e3b_length_only = True
        return ({"radial_embedding": emb},
                {"radial_embedding": (attrs["input"][0], self.irreps_out["radial_embedding"])})


class Broadcast(Module):
    """graph features -> nodes or edges"""

    def __init__(self, irreps_in, irreps_out, to):
        super().__init__()
        self.init_irreps(input=irreps_in, output=irreps_out, output_keys=["output"])
        if to not in ("node", "edge"):
            raise ValueError(to)
        self.to = to

    def forward(self, data, attrs):
        assert attrs["input"][0] == "graph"
        seg = data["_node_segment" if self.to == "node" else "_edge_segment"]
        return {"output": data["input"][seg.to(data["input"].device)]}, {"output": (self.to, self.irreps_out["output"])}


class OneHotEncoding(Module):
    def __init__(self, num_types, irreps_out, irreps_in="0x0e"):
        super().__init__()
        self.num_types = num_types
        self.init_irreps(input=irreps_in, one_hot=irreps_out, output_keys="one_hot")

    def forward(self, data, attrs):
        idx = data["input"].squeeze(-1)
        # the reference hard-codes float32 (D6); follow the default dtype so the fp64 mode works
        one_hot = torch.nn.functional.one_hot(idx, num_classes=self.num_types).to(torch.get_default_dtype())
        one_hot._e3b_onehot = (idx, self.num_types)     # provenance: whatever is computed row by row from this is a
                                                        # function of the species (see PointwiseLinear)
        return {"one_hot": one_hot}, {"one_hot": (attrs["input"][0], self.irreps_out["one_hot"])}


class RelativePositionEncoding(Module):
    """radial embedding of the sequence separation of an edge's endpoints (same segment only)"""

    def __init__(self, radial_encoding, segment, irreps_out, id=None):
        super().__init__()
        self.init_irreps(input=segment, output=irreps_out, id=id, output_keys=["output"])
        cfg = dict(radial_encoding.items())
        cfg["irreps_in"], cfg["irreps_out"] = "1x0e", self.irreps_out["output"]
        self.radial = build(cfg)

    def forward(self, data, attrs):
        seg, (src, dst) = data["input"], data["edge_index"]
        ids = data["id"] if self.irreps_in.get("id") is not None else None
        rel = (ids[src] - ids[dst]) if ids is not None else (src - dst)
        same = (seg[src] == seg[dst]).float()
        rel = same * rel.view(-1, 1) + (1 - same) * 1e5
        out, _ = self.radial({"input": rel.to(torch.get_default_dtype())}, attrs)
        return {"output": out["radial_embedding"]}, {"output": ("edge", self.irreps_out["output"])}
