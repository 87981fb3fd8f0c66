"""Second, independent derivation of the real Wigner-3j tensors (oracle; test infrastructure only).

``oracle/wigner.py`` builds the tensors analytically (SU(2) Clebsch-Gordan by the Racah formula,
rotated into the real basis -- the e3nn >= 0.5 recipe).  e3nn 0.4.4, which the reference pins
(``/root/reference/requirements.txt:27``), instead ships a table that was generated numerically:
the one-dimensional null space of ``D^{l1}(R) (x) D^{l2}(R) (x) D^{l3}(R) - 1`` over a handful of
rotations R, normalised to Frobenius norm 1, with the global sign fixed by a rule on the entries.
This module restates that construction using NOTHING from ``oracle/wigner.py`` except the
spherical-harmonics polynomials (which define the basis): the representation matrices come from a
least-squares fit of ``Y^l(R r) = D^l(R) Y^l(r)`` on random points, not from any 3j tensor.

For every triple it reports
  * whether the null space is one-dimensional and equal, up to sign, to the analytic tensor,
  * the sign the analytic tensor has under each candidate sign rule of the generated table:
      rule "first"  -- first non-zero entry in flat (C) order of the l1<=l2<=l3 representative > 0;
      rule "centre" -- entry [l1, l2, l3] (m = 0,0,0) > 0 when non-zero, else rule "first";
  * for (1, l-1, l): the sign that the spherical-harmonics polynomials themselves force (e3nn
    generates Y^l from Y^1 (x) Y^{l-1} with that tensor and a positive normalisation, so the
    published polynomials -- Y^l_0 > 0 at the +y pole -- pin these signs without any table).

A sign a rule would flip relative to the analytic tensor is "ambiguous": only genuine e3nn 0.4.4 can
settle it (``tools/check_against_e3nn.py``).  ``tests/test_oracle_w3j_nullspace.py`` records the result.
"""
import itertools
import math

import numpy as np
import torch

from .wigner import spherical_harmonics_raw


def _sh(l, pts):
    """[P, 2l+1] float64: the (un-normalised) real SH polynomials of degree l at points pts [P, 3]"""
    t = torch.as_tensor(pts, dtype=torch.float64)
    comps = spherical_harmonics_raw(l, t[:, 0], t[:, 1], t[:, 2])
    return torch.stack(comps, dim=-1).numpy()


def _rotation(rng):
    q, r = np.linalg.qr(rng.standard_normal((3, 3)))
    q = q * np.sign(np.diag(r))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def rep_matrix(l, R, rng):
    """D^l(R) in the basis of the SH polynomials: least squares over random unit vectors of
    Y^l(R r) = D Y^l(r).  Exact (residual ~1e-15) because Y^l spans an invariant subspace."""
    pts = rng.standard_normal((8 * (2 * l + 1) + 8, 3))
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    A = _sh(l, pts)                      # [P, d]
    B = _sh(l, pts @ R.T)                # [P, d]  rows Y(R r)
    X, res, rank, _ = np.linalg.lstsq(A, B, rcond=None)   # A X = B  ->  D = X^T
    assert rank == 2 * l + 1
    D = X.T
    assert np.abs(A @ X - B).max() < 1e-12
    return D


def nullspace_w3j(l1, l2, l3, n_rot=5, seed=0):
    """-> (Q [2l1+1, 2l2+1, 2l3+1] with Frobenius norm 1 and arbitrary sign, gap) where gap is the second
    smallest eigenvalue of sum_R (D(x)D(x)D - 1)^T (D(x)D(x)D - 1) (> 0 <=> the invariant is unique)."""
    rng = np.random.default_rng(seed + 1000 * (l1 * 16 + l2 * 4 + l3))
    d1, d2, d3 = 2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1
    n = d1 * d2 * d3
    B = np.zeros((n, n))
    for _ in range(n_rot):
        R = _rotation(rng)
        D1, D2, D3 = rep_matrix(l1, R, rng), rep_matrix(l2, R, rng), rep_matrix(l3, R, rng)
        M = np.einsum("il,jm,kn->ijklmn", D1, D2, D3).reshape(n, n) - np.eye(n)
        B += M.T @ M
    w, v = np.linalg.eigh(B)
    assert w[0] < 1e-10, (l1, l2, l3, w[0])
    Q = v[:, 0].reshape(d1, d2, d3)
    Q = Q / np.linalg.norm(Q)
    Q[np.abs(Q) < 1e-12] = 0.0
    return Q, float(w[1]) if n > 1 else float("inf")


def sign_first(Q):
    flat = Q.reshape(-1)
    nz = flat[np.abs(flat) > 1e-10 * np.abs(flat).max()]
    return 1.0 if nz[0] > 0 else -1.0


def sign_centre(Q, l1, l2, l3):
    c = Q[l1, l2, l3]
    if abs(c) > 1e-10:
        return 1.0 if c > 0 else -1.0
    return sign_first(Q)


def sh_recursion_sign(l, Q):
    """Sign s such that  s * sum_ij Q[i,j,k] Y^1_i Y^{l-1}_j  is a POSITIVE multiple of the published Y^l_k
    (Q = a (1, l-1, l) tensor).  None if Q is not of that shape."""
    rng = np.random.default_rng(7)
    pts = rng.standard_normal((32, 3))
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    lhs = np.einsum("ijk,pi,pj->pk", Q, _sh(1, pts), _sh(l - 1, pts))
    rhs = _sh(l, pts)
    ratio = (lhs * rhs).sum() / (rhs * rhs).sum()
    assert np.abs(lhs - ratio * rhs).max() < 1e-10, "Y^1 (x) Y^{l-1} -> l is not proportional to Y^l"
    return 1.0 if ratio > 0 else -1.0


def triples(lmax=3):
    for l1, l2, l3 in itertools.product(range(lmax + 1), repeat=3):
        if l1 <= l2 <= l3 and abs(l1 - l2) <= l3 <= l1 + l2:
            yield l1, l2, l3


def report(lmax=3):
    """One record per l1<=l2<=l3 representative (other orders follow by the permutation rule of
    SURVEY A.3, which holds exactly for the analytic tensors -- tests/test_oracle_kat.py)."""
    from . import wigner

    out = []
    for (l1, l2, l3) in triples(lmax):
        A = wigner._w3j_analytic(l1, l2, l3)
        Q, gap = nullspace_w3j(l1, l2, l3)
        dot = float((A * Q).sum())
        rec = {
            "triple": [l1, l2, l3],
            "parity_of_sum": (l1 + l2 + l3) % 2,
            "nullspace_dim_is_1": bool(gap > 1e-8),
            "equal_up_to_sign": bool(abs(abs(dot) - 1) < 1e-10 and np.abs(A - np.sign(dot) * Q).max() < 1e-10),
            "analytic_sign_under_rule_first": sign_first(A),
            "analytic_sign_under_rule_centre": sign_centre(A, l1, l2, l3),
            "analytic_sign_forced_by_sh_polynomials": sh_recursion_sign(l3, A) if (l1 == 1 and l2 == l3 - 1) else None,
        }
        forced = rec["analytic_sign_forced_by_sh_polynomials"]
        rec["pinned"] = bool(forced == 1.0 or (rec["analytic_sign_under_rule_first"] == 1.0
                                               and rec["analytic_sign_under_rule_centre"] == 1.0))
        out.append(rec)
    return out


if __name__ == "__main__":
    import json

    print(json.dumps(report(), indent=1))
