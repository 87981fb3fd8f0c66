"""The oracle's real Wigner-3j tensors checked by a second, independent derivation (null space of
D (x) D (x) D - 1, the way e3nn 0.4.4's table was generated; oracle/wigner_nullspace.py), and the exact list of
triples whose GLOBAL SIGN cannot be settled without genuine e3nn 0.4.4."""
import json
import os
import subprocess
import sys

import numpy as np

from oracle import wigner, wigner_nullspace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_nullspace_equals_analytic_up_to_sign():
    recs = wigner_nullspace.report(lmax=3)
    assert len(recs) == 13
    for r in recs:
        assert r["nullspace_dim_is_1"], r
        assert r["equal_up_to_sign"], r


def test_sign_ambiguity_is_exactly_this_list():
    """Even l1+l2+l3 with a (1, l-1, l) shape: the published SH polynomials force the analytic sign.  What a sign
    rule on the generated table could still flip (relative to the analytic tensor the product and oracle use):"""
    recs = {tuple(r["triple"]): r for r in wigner_nullspace.report(lmax=3)}
    forced = {t for t, r in recs.items() if r["analytic_sign_forced_by_sh_polynomials"] is not None}
    assert forced == {(1, 1, 2), (1, 2, 3)}
    assert all(recs[t]["analytic_sign_forced_by_sh_polynomials"] == 1.0 for t in forced)
    # rule "first non-zero positive" would flip (1,1,2) and (1,2,3), which the SH polynomials forbid -> that rule
    # cannot be the one behind the table the SH code was generated from; rule "centre, else first" is consistent
    assert recs[(1, 1, 2)]["analytic_sign_under_rule_first"] == -1.0
    assert all(recs[t]["analytic_sign_under_rule_centre"] == 1.0 for t in forced)
    flips_centre = sorted(t for t, r in recs.items() if r["analytic_sign_under_rule_centre"] < 0)
    flips_first = sorted(t for t, r in recs.items() if r["analytic_sign_under_rule_first"] < 0)
    assert flips_centre == [(1, 2, 2), (1, 3, 3)]
    assert flips_first == [(1, 1, 2), (1, 2, 2), (1, 2, 3), (1, 3, 3), (2, 2, 2), (2, 3, 3)]
    # the product's preset table (csrc/cg_tables.cuh kCgSign044, oracle preset "e3nn044") encodes rule "centre"
    for t in recs:
        assert wigner._sign_e3nn044(*t) == recs[t]["analytic_sign_under_rule_centre"], t


def test_permutation_rule_holds_for_analytic_tensors():
    """0.4.4 stores l1<=l2<=l3 only and maps other orders by: even permutation -> transpose, odd -> transpose
    and multiply by (-1)^(l1+l2+l3).  The analytic tensors satisfy it exactly, so one sign per
    sorted triple is the whole ambiguity."""
    for (l1, l2, l3) in wigner_nullspace.triples(3):
        base = wigner._w3j_analytic(l1, l2, l3)
        s = (-1.0) ** (l1 + l2 + l3)
        assert np.allclose(wigner._w3j_analytic(l1, l3, l2), s * base.transpose(0, 2, 1), atol=1e-12)
        assert np.allclose(wigner._w3j_analytic(l2, l1, l3), s * base.transpose(1, 0, 2), atol=1e-12)
        assert np.allclose(wigner._w3j_analytic(l3, l2, l1), s * base.transpose(2, 1, 0), atol=1e-12)
        assert np.allclose(wigner._w3j_analytic(l2, l3, l1), base.transpose(1, 2, 0), atol=1e-12)
        assert np.allclose(wigner._w3j_analytic(l3, l1, l2), base.transpose(2, 0, 1), atol=1e-12)


def test_convention_dump_is_current():
    """tests/golden/conventions_oracle.json is what `tools/check_against_e3nn.py --against` diffs genuine e3nn
    with; it must be the oracle's current tables."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_against_e3nn.py"), "--side", "oracle",
                          "--against", os.path.join(ROOT, "tests", "golden", "conventions_oracle.json")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "mismatches: 0" in out.stdout
    with open(os.path.join(ROOT, "tests", "golden", "conventions_oracle.json")) as f:
        d = json.load(f)
    assert len(d["wigner_3j"]) == sum(1 for a in range(4) for b in range(4) for c in range(abs(a - b), min(3, a + b) + 1))
