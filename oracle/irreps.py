"""Irreps algebra of e3nn 0.4.4, restated (oracle; test infrastructure only).

Follows SURVEY.md Appendix A.1.  Call sites in the reference that rely on it:
``e3_layers/utils/utils.py:87-96`` (tp_path_exists), ``configs/layer_configs.py:32-37,86-96``,
``nn/message_passing.py:167-207``, ``nn/pointwise.py:61-76`` (``Irreps.sort``),
``data/data.py:84`` / ``data/batch.py:102`` (``Irreps(...).dim``).
"""
import collections
import re
from typing import List, Tuple


class Irrep(tuple):
    """(l, p) with p = +1 ('e') or -1 ('o').  A plain tuple subclass, so ordering is the
    tuple ordering: l first, then p with -1 (odd) < +1 (even)  [SURVEY A.1]."""

    def __new__(cls, l, p=None):
        if p is None:
            if isinstance(l, Irrep):
                return l
            if isinstance(l, str):
                m = re.fullmatch(r"\s*(\d+)([eoy])\s*", l)
                if m is None:
                    raise ValueError(f"cannot parse irrep {l!r}")
                ll = int(m.group(1))
                p = {"e": 1, "o": -1, "y": (-1) ** ll}[m.group(2)]
                l = ll
            elif isinstance(l, tuple):
                l, p = l
        if not isinstance(l, int) or l < 0:
            raise ValueError(f"l must be a non-negative int, got {l}")
        if p not in (-1, 1):
            raise ValueError(f"parity must be +-1, got {p}")
        return super().__new__(cls, (l, p))

    @property
    def l(self):
        return self[0]

    @property
    def p(self):
        return self[1]

    @property
    def dim(self):
        return 2 * self.l + 1

    def is_scalar(self):
        return self.l == 0 and self.p == 1

    def __repr__(self):
        return f"{self.l}{'e' if self.p == 1 else 'o'}"

    def __mul__(self, other):
        other = Irrep(other)
        p = self.p * other.p
        return [Irrep(l, p) for l in range(abs(self.l - other.l), self.l + other.l + 1)]

    def __rmul__(self, mul):
        assert isinstance(mul, int)
        return Irreps([(mul, self)])

    def __add__(self, other):
        return Irreps(self) + Irreps(other)


class _MulIr(tuple):
    def __new__(cls, mul, ir):
        return super().__new__(cls, (int(mul), Irrep(ir)))

    @property
    def mul(self):
        return self[0]

    @property
    def ir(self):
        return self[1]

    @property
    def dim(self):
        return self.mul * self.ir.dim

    def __repr__(self):
        return f"{self.mul}x{self.ir}"


class Irreps(tuple):
    """Ordered list of (mul, Irrep); order preserved as written."""

    def __new__(cls, irreps=None):
        if isinstance(irreps, Irreps):
            return super().__new__(cls, irreps)
        out = []
        if irreps is None:
            pass
        elif isinstance(irreps, Irrep):
            out.append(_MulIr(1, irreps))
        elif isinstance(irreps, str):
            if irreps.strip() != "":
                for term in irreps.split("+"):
                    term = term.strip()
                    if "x" in term:
                        mul, ir = term.split("x")
                        out.append(_MulIr(int(mul), Irrep(ir)))
                    else:
                        out.append(_MulIr(1, Irrep(term)))
        else:
            for item in irreps:
                if isinstance(item, _MulIr):
                    out.append(item)
                elif isinstance(item, Irrep):
                    out.append(_MulIr(1, item))
                elif isinstance(item, str):
                    out.append(_MulIr(1, Irrep(item)))
                else:
                    mul, ir = item
                    out.append(_MulIr(mul, Irrep(ir)))
        return super().__new__(cls, out)

    @staticmethod
    def spherical_harmonics(lmax, p=-1):
        return Irreps([(1, (l, p ** l)) for l in range(lmax + 1)])

    def slices(self):
        s, i = [], 0
        for mul_ir in self:
            s.append(slice(i, i + mul_ir.dim))
            i += mul_ir.dim
        return s

    @property
    def dim(self):
        return sum(mi.dim for mi in self)

    @property
    def num_irreps(self):
        return sum(mi.mul for mi in self)

    @property
    def ls(self):
        return [mi.ir.l for mi in self for _ in range(mi.mul)]

    @property
    def lmax(self):
        if len(self) == 0:
            raise ValueError("empty irreps has no lmax")
        return max(mi.ir.l for mi in self)

    def simplify(self):
        """Merge ADJACENT equal irreps, drop mul 0."""
        out = []
        for mul, ir in self:
            if out and out[-1][1] == ir:
                out[-1] = (out[-1][0] + mul, ir)
            elif mul > 0:
                out.append((mul, ir))
        return Irreps(out)

    def remove_zero_multiplicities(self):
        return Irreps([(mul, ir) for mul, ir in self if mul > 0])

    def sort(self):
        """Stable sort of blocks by Irrep.  Returns (irreps, p, inv) with p[old] = new."""
        Ret = collections.namedtuple("sort", ["irreps", "p", "inv"])
        out = sorted([(ir, i, mul) for i, (mul, ir) in enumerate(self)])
        inv = tuple(i for _, i, _ in out)
        p = [0] * len(inv)
        for new, old in enumerate(inv):
            p[old] = new
        irreps = Irreps([(mul, ir) for ir, _, mul in out])
        return Ret(irreps, tuple(p), inv)

    def count(self, ir):
        ir = Irrep(ir)
        return sum(mul for mul, jr in self if jr == ir)

    def __contains__(self, ir):
        try:
            ir = Irrep(ir)
        except Exception:
            return False
        return any(jr == ir for _, jr in self)

    def __add__(self, other):
        return Irreps(tuple.__add__(self, Irreps(other)))

    def __mul__(self, n):
        if isinstance(n, int):
            return Irreps(tuple.__mul__(self, n))
        return NotImplemented

    __rmul__ = __mul__

    def __getitem__(self, i):
        x = tuple.__getitem__(self, i)
        if isinstance(i, slice):
            return Irreps(x)
        return x

    def __eq__(self, other):
        try:
            other = Irreps(other)
        except Exception:
            return False
        return tuple.__eq__(self, other)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return tuple.__hash__(self)

    def __repr__(self):
        return "+".join(f"{mi}" for mi in self)

    # -- representation matrices (only what equivariance tests need) --------------------
    def D_from_matrix(self, R):
        """Block-diagonal real representation of O(3) element R (3x3, det +-1).

        Built from the oracle's own spherical harmonics so that it is, by construction,
        the representation the SH / wigner_3j of this package transform under."""
        import torch
        from .wigner import wigner_D_from_matrix

        blocks = []
        for mul, ir in self:
            d = wigner_D_from_matrix(ir.l, ir.p, R)
            for _ in range(mul):
                blocks.append(d)
        if not blocks:
            return torch.zeros(0, 0, dtype=R.dtype)
        return torch.block_diag(*blocks)
