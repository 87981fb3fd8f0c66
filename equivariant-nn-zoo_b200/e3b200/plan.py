"""Host-side description of one tensor-product convolution (the "plan"): which (input block,
SH block, l_out) paths exist, where each path's weights sit in the per-edge weight row and
where its output block sits in the sorted intermediate irreps.

Mirrors the instruction builder of ``TensorProductExpansion.__init__``
(``e3_layers/nn/pointwise.py:61-76``): left block slow, right block, then every l_out of the
product that occurs in the requested output irreps; output blocks sorted by irrep with the
stable order of creation; weights in creation order (e3nn 'uvu', mul_in2 = 1 -> `mul` numbers
per path)."""
from collections import namedtuple

from .irreps import Irreps

Path = namedtuple("Path", ["i_in", "i_sh", "ir_out", "slot"])


class TPStructure:
    def __init__(self, irreps_in, irreps_sh, irreps_out):
        self.irreps_in = Irreps(irreps_in)
        self.irreps_sh = Irreps(irreps_sh)
        self.irreps_out = Irreps(irreps_out)
        raw, mid = [], []
        for i, (mul, ir1) in enumerate(self.irreps_in):
            for j, (mul2, ir2) in enumerate(self.irreps_sh):
                for ir_out in ir1 * ir2:
                    if ir_out in self.irreps_out:
                        raw.append((i, j, ir_out))
                        mid.append((mul, ir_out))
        srt = Irreps(mid).sort()
        self.irreps_mid = srt.irreps
        self.paths = [Path(i, j, ir, srt.p[k]) for k, (i, j, ir) in enumerate(raw)]
        self.weight_numel = sum(self.irreps_in[p.i_in].mul * self.irreps_sh[p.i_sh].mul for p in self.paths)

    @property
    def uniform_mul(self):
        muls = {b.mul for b in self.irreps_in}
        return muls.pop() if len(muls) == 1 and all(b.mul == 1 for b in self.irreps_sh) else None

    def key(self):
        """Arithmetic signature (independent of mul and of parities)."""
        base, kstride, _ = self.y_layout()
        return (tuple(b.ir.l for b in self.irreps_in), tuple(b.ir.l for b in self.irreps_sh),
                tuple((p.i_in, p.i_sh, p.ir_out.l, p.slot, base[p.slot], kstride[p.slot]) for p in self.paths))

    # layouts in units of `mul` scalars (imu layout: [component][channel])
    def x_comp_offsets(self):
        out, o = [], 0
        for b in self.irreps_in:
            out.append(o)
            o += b.ir.dim
        return out, o

    def sh_comp_offsets(self):
        out, o = [], 0
        for b in self.irreps_sh:
            out.append(o)
            o += b.ir.dim
        return out, o

    def y_layout(self):
        """Output row layout, in units of `mul` scalars.  Slots (sorted irreps_mid blocks) with
        the same irrep form one group laid out [k][slot-in-group][u] -- exactly the imu layout
        of the SIMPLIFIED irreps_mid block the post-linear consumes.
        -> (base[slot], kstride[slot], total components)"""
        base, kstride = [0] * len(self.irreps_mid), [0] * len(self.irreps_mid)
        o, s = 0, 0
        while s < len(self.irreps_mid):
            e = s
            while e < len(self.irreps_mid) and self.irreps_mid[e].ir == self.irreps_mid[s].ir:
                e += 1
            n = e - s
            for t in range(s, e):
                base[t], kstride[t] = o + (t - s), n
            o += n * self.irreps_mid[s].ir.dim
            s = e
        return base, kstride, o


def reference_structures(l_max_features, sh="1x0e+1x1o+1x2e", n_layers=6):
    """The distinct TP structures the reference's featureModel produces
    (``configs/layer_configs.py:86-98`` + ``nn/message_passing.py:171-207``): layer i maps the
    reachable feature irreps to the next reachable set; the conv output adds 0e gate scalars."""
    sh = Irreps(sh)
    full = Irreps("+".join(f"1x{n}e+1x{n}o" for n in range(l_max_features + 1)))

    def reachable(cur, ir):
        return any(ir in a.ir * b.ir for a in cur for b in sh)

    cur = Irreps("1x0e")
    out = []
    for _ in range(n_layers):
        nxt = Irreps([(1, b.ir) for b in full if reachable(cur, b.ir)])
        conv_out = nxt + Irreps("1x0e")
        st = TPStructure(cur, sh, conv_out)
        if all(st.key() != s.key() for s in out):
            out.append(st)
        cur = nxt
    return out


def scalar_output_restriction(st):
    """`st` restricted to the paths that produce 0e: what the backward pass of the LAST interaction block needs when the
    only consumer of its output is a head that reads the 0e scalars (an energy read-out: reference addEnergyOutput,
    configs/layer_configs.py:101-113) -- the gradient of every other output block is exactly zero there."""
    return TPStructure(st.irreps_in, st.irreps_sh, Irreps([(st.irreps_in[0].mul, "0e")]))


def output_restriction(st, irs):
    """`st` restricted to the paths whose output irrep is one of `irs` (any order)"""
    keep = [b for b in st.irreps_out if any(b.ir == ir for ir in irs)]
    seen, uniq = set(), []
    for b in keep:
        if str(b.ir) not in seen:
            seen.add(str(b.ir))
            uniq.append((st.irreps_in[0].mul, b.ir))
    return TPStructure(st.irreps_in, st.irreps_sh, Irreps(uniq))


def generated_structures():
    """The list (in order) of structures csrc/gen_tp.py emits unrolled kernels for: the reference's layer structures for
    l_max 2 and 3, then their restrictions to the 0e output (appended, so the numbering of the former is stable)."""
    out = []
    for lm in (2, 3):
        for st in reference_structures(lm):
            if all(st.key() != s.key() for s in out):
                out.append(st)
    base = list(out)
    for st in base:
        pr = scalar_output_restriction(st)
        if pr.paths and all(pr.key() != s.key() for s in out):
            out.append(pr)
    # one block earlier the live outputs are the inputs of those paths: 0e, 1o, 2e (interaction.FusedInteraction.restricted)
    nat = Irreps("1x0e+1x1o+1x2e")
    for st in base:
        pr = output_restriction(st, [b.ir for b in nat])
        if pr.paths and len(pr.paths) < len(st.paths) and all(pr.key() != s.key() for s in out):
            out.append(pr)
    return out


def with_mul(st, mul):
    """Same structure with every input block at multiplicity `mul`."""
    return TPStructure(Irreps([(mul, b.ir) for b in st.irreps_in]), st.irreps_sh,
                       Irreps([(mul, b.ir) for b in st.irreps_out]))
