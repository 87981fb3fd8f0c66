"""Irreducible-representation bookkeeping for the drop-in modules (replaces ``e3nn.o3.Irrep`` /
``e3nn.o3.Irreps`` for everything the reference asks of them: parsing ``"64x0e+64x1o"``
strings, ``dim``, ``slices``, ``simplify``, ``sort``, products and membership; reference uses
at ``utils/utils.py:87-96``, ``configs/layer_configs.py:32-37,86-96``,
``nn/message_passing.py:167-207``, ``nn/pointwise.py:61-76``, ``data/data.py:84``).

Blocks keep the order in which they were written.  Sorting orders by (l, parity) with odd
before even, i.e. 0o < 0e < 1o < 1e < ... (tuple order of (l, p), p = -1 | +1)."""
from collections import namedtuple

_PCHAR = {1: "e", -1: "o"}


class Irrep(namedtuple("Irrep", ["l", "p"])):
    __slots__ = ()

    def __new__(cls, l, p=None):
        if p is None:
            if isinstance(l, Irrep):
                return l
            if isinstance(l, str):
                s = l.strip()
                l, c = int(s[:-1]), s[-1]
                p = {"e": 1, "o": -1, "y": (-1) ** l}[c]
            else:
                l, p = l
        assert isinstance(l, int) and l >= 0 and p in (1, -1), (l, p)
        return super().__new__(cls, l, p)

    @property
    def dim(self):
        return 2 * self.l + 1

    def is_scalar(self):
        return self.l == 0 and self.p == 1

    def __mul__(self, other):
        other = Irrep(other)
        return [Irrep(l, self.p * other.p) for l in range(abs(self.l - other.l), self.l + other.l + 1)]

    def __repr__(self):
        return f"{self.l}{_PCHAR[self.p]}"


class MulIr(namedtuple("MulIr", ["mul", "ir"])):
    __slots__ = ()

    @property
    def dim(self):
        return self.mul * self.ir.dim

    def __repr__(self):
        return f"{self.mul}x{self.ir}"


def _parse(spec):
    if spec is None:
        return []
    if isinstance(spec, Irrep):
        return [MulIr(1, spec)]
    if isinstance(spec, str):
        out = []
        for tok in filter(None, (t.strip() for t in spec.split("+"))):
            mul, _, ir = tok.rpartition("x")
            out.append(MulIr(int(mul) if mul else 1, Irrep(ir)))
        return out
    out = []
    for item in spec:
        if isinstance(item, MulIr):
            out.append(item)
        elif isinstance(item, (Irrep, str)):
            out.append(MulIr(1, Irrep(item)))
        else:
            mul, ir = item
            out.append(MulIr(int(mul), Irrep(ir)))
    return out


class Irreps(tuple):
    def __new__(cls, spec=None):
        if isinstance(spec, Irreps):
            return spec
        return super().__new__(cls, _parse(spec))

    # sizes --------------------------------------------------------------------------------
    @property
    def dim(self):
        return sum(b.dim for b in self)

    @property
    def num_irreps(self):
        return sum(b.mul for b in self)

    @property
    def lmax(self):
        return max(b.ir.l for b in self)

    def slices(self):
        out, start = [], 0
        for b in self:
            out.append(slice(start, start + b.dim))
            start += b.dim
        return out

    def offsets(self):
        return [s.start for s in self.slices()]

    # transforms ---------------------------------------------------------------------------
    def simplify(self):
        merged = []
        for mul, ir in self:
            if mul == 0:
                continue
            if merged and merged[-1][1] == ir:
                merged[-1] = (merged[-1][0] + mul, ir)
            else:
                merged.append((mul, ir))
        return Irreps(merged)

    def sort(self):
        """-> (sorted irreps, p, inv): block `old` moves to position p[old]; inv[new] = old."""
        order = sorted(range(len(self)), key=lambda i: (tuple(self[i].ir), i))
        p = [0] * len(self)
        for new, old in enumerate(order):
            p[old] = new
        Sorted = namedtuple("Sorted", ["irreps", "p", "inv"])
        return Sorted(Irreps([self[i] for i in order]), tuple(p), tuple(order))

    def count(self, ir):
        ir = Irrep(ir)
        return sum(b.mul for b in self if b.ir == ir)

    def __contains__(self, ir):
        try:
            ir = Irrep(ir)
        except Exception:
            return False
        return any(b.ir == ir for b in self)

    def __add__(self, other):
        return Irreps(list(self) + list(Irreps(other)))

    def __getitem__(self, idx):
        got = tuple.__getitem__(self, idx)
        return Irreps(got) if isinstance(idx, slice) else got

    def __eq__(self, other):
        try:
            return tuple.__eq__(self, Irreps(other))
        except Exception:
            return False

    def __ne__(self, other):
        return not self == other

    __hash__ = tuple.__hash__

    def __repr__(self):
        return "+".join(repr(b) for b in self)

    @staticmethod
    def spherical_harmonics(lmax):
        return Irreps([(1, (l, (-1) ** l)) for l in range(lmax + 1)])
