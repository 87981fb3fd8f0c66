#!/usr/bin/env python3
"""Pins every convention the oracle restates against GENUINE e3nn (pin: e3nn==0.4.4,
/root/reference/requirements.txt:27) in one command.

    python tools/check_against_e3nn.py                      # on a machine WITH e3nn: diff live, exit 1 on mismatch
    python tools/check_against_e3nn.py --dump out.json      # write the convention tables of whichever side is importable
    python tools/check_against_e3nn.py --side oracle --dump tests/golden/conventions_oracle.json   (committed)
    python tools/check_against_e3nn.py --against tests/golden/conventions_oracle.json               (with e3nn: diff vs the file)

The build container has no e3nn (oracle/__init__.py: "parity unpinned"), so the oracle side of the dump is
committed; anyone with e3nn 0.4.4 runs the last form and gets a per-entry verdict.  Dumped: real wigner_3j for
every triple l <= 3 (all orders), component-normalised spherical harmonics l <= 3 on seeded probes, the
normalize2mom constants, and seeded outputs of o3.Linear, the externally weighted 'uvu' TensorProduct,
FullyConnectedTensorProduct, nn.Gate and nn.FullyConnectedNet with weights copied by name."""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

torch.set_default_dtype(torch.float64)

ACTS = {"ssp": lambda x: torch.nn.functional.softplus(x) - math.log(2.0), "silu": torch.nn.functional.silu,
        "tanh": torch.tanh, "abs": torch.abs, "tanhlu": lambda x: torch.tanh(x) * x.abs()}
IRR = "8x0e+8x0o+8x1e+8x1o+8x2e+8x2o"
SH = "1x0e+1x1o+1x2e"


def _seeded(shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def _fill(module, seed):
    """identical parameter values on both sides: every parameter, in name order, from one seeded stream"""
    g = torch.Generator().manual_seed(seed)
    for name, p in sorted(module.named_parameters()):
        p.data.copy_(torch.randn(p.shape, generator=g))


def tables(side):
    if side == "e3nn":
        import e3nn
        from e3nn import nn as e_nn, o3
        from e3nn.math import normalize2mom

        w3j, sph = o3.wigner_3j, o3.spherical_harmonics
        Linear, TP, FCTP, Gate, FCN, Irreps = (o3.Linear, o3.TensorProduct, o3.FullyConnectedTensorProduct, e_nn.Gate,
                                               e_nn.FullyConnectedNet, o3.Irreps)
        n2m = lambda f: float(normalize2mom(f).cst)
        version = e3nn.__version__
    else:
        from oracle import e3nn_ops, wigner
        from oracle.irreps import Irreps

        w3j, sph = wigner.wigner_3j, wigner.spherical_harmonics
        Linear, TP, FCTP, Gate, FCN = (e3nn_ops.Linear, e3nn_ops.TensorProduct, e3nn_ops.FullyConnectedTensorProduct,
                                       e3nn_ops.Gate, e3nn_ops.FullyConnectedNet)
        n2m = e3nn_ops.normalize2mom_constant
        version = "oracle restatement (sign preset %s)" % wigner.SIGN_PRESET
    out = {"side": side, "version": version, "wigner_3j": {}, "normalize2mom": {}, "ops": {}}
    for l1 in range(4):
        for l2 in range(4):
            for l3 in range(abs(l1 - l2), min(3, l1 + l2) + 1):
                out["wigner_3j"][f"{l1},{l2},{l3}"] = w3j(l1, l2, l3).reshape(-1).tolist()
    v = _seeded((16, 3), 0)
    out["spherical_harmonics_l0123_component"] = sph([0, 1, 2, 3], v, True, "component").reshape(-1).tolist()
    for name, f in ACTS.items():
        out["normalize2mom"][name] = n2m(f)
    x, y, z, r = _seeded((6, Irreps(IRR).dim), 1), _seeded((6, 9), 2), _seeded((6, 4), 3), _seeded((6, 8), 4)
    m = Linear(IRR, IRR)
    _fill(m, 10)
    out["ops"]["Linear"] = m(x).reshape(-1).tolist()
    irr, sh = Irreps(IRR), Irreps(SH)
    ins = [(i, j, k, "uvu", True) for i in range(6) for j in range(3) for k in range(6) if irr[k].ir in irr[i].ir * sh[j].ir]
    m = TP(IRR, SH, IRR, ins, shared_weights=False, internal_weights=False)
    out["ops"]["TensorProduct_uvu"] = m(x, y, _seeded((6, m.weight_numel), 5)).reshape(-1).tolist()
    m = FCTP(IRR, "4x0e", IRR)
    _fill(m, 11)
    out["ops"]["FullyConnectedTensorProduct"] = m(x, z).reshape(-1).tolist()
    m = FCN([8, 16, 16, 4], ACTS["ssp"])
    _fill(m, 12)
    out["ops"]["FullyConnectedNet_ssp"] = m(r).reshape(-1).tolist()
    gate = Gate("8x0e+8x0o", [ACTS["silu"], ACTS["tanhlu"]], "32x0e", [ACTS["silu"]], "8x1e+8x1o+8x2e+8x2o")
    out["ops"]["Gate"] = gate(_seeded((6, gate.irreps_in.dim), 6)).reshape(-1).tolist()
    return out


def diff(a, b, tol=1e-10):
    """-> list of (name, max abs diff, ok[, note]) over every entry present on both sides"""
    rows = []
    for key, va in a["wigner_3j"].items():
        ta, tb = torch.tensor(va), torch.tensor(b["wigner_3j"][key])
        d = float((ta - tb).abs().max())
        note = ""
        if d > tol and float((ta + tb).abs().max()) < tol:
            note = "GLOBAL SIGN FLIP (convention risk R1)"
        rows.append((f"wigner_3j({key})", d, d < tol, note))
    k = "spherical_harmonics_l0123_component"
    rows.append((k, float((torch.tensor(a[k]) - torch.tensor(b[k])).abs().max()), None, ""))
    for name in a["normalize2mom"]:
        rows.append((f"normalize2mom[{name}]", abs(a["normalize2mom"][name] - b["normalize2mom"][name]), None, ""))
    for name in a["ops"]:
        rows.append((f"op {name}", float((torch.tensor(a["ops"][name]) - torch.tensor(b["ops"][name])).abs().max()), None, ""))
    out = []
    for name, d, ok, note in rows:
        if ok is None:
            ok = d < (1e-6 if name.startswith("normalize2mom") or "FullyConnectedNet" in name or "Gate" in name else tol)
        out.append((name, d, ok, note))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", choices=["auto", "oracle", "e3nn"], default="auto")
    ap.add_argument("--dump")
    ap.add_argument("--against")
    a = ap.parse_args()
    have = True
    try:
        import e3nn  # noqa: F401
    except ImportError:
        have = False
    side = a.side if a.side != "auto" else ("e3nn" if have else "oracle")
    if side == "e3nn" and not have:
        print("e3nn is not importable here")
        return 2
    mine = tables(side)
    if a.dump:
        with open(a.dump, "w") as f:
            json.dump(mine, f)
        print(f"wrote {a.dump} ({side}: {mine['version']})")
    other = None
    if a.against:
        with open(a.against) as f:
            other = json.load(f)
    elif side == "e3nn" and a.side == "auto":
        other = tables("oracle")
    if other is None:
        if not a.dump:
            print("e3nn is not importable here; nothing compared (use --dump / --against)")
            return 2
        return 0
    print(f"comparing {mine['side']} ({mine['version']}) with {other['side']} ({other['version']})")
    bad = 0
    for name, d, ok, note in diff(mine, other):
        bad += 0 if ok else 1
        print(f"{'ok  ' if ok else 'FAIL'} {name}: max abs diff {d:.3e} {note}")
    print("mismatches:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
