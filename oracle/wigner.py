"""Real Wigner-3j tensors and real spherical harmonics in the e3nn basis (oracle; test
infrastructure only).  Follows SURVEY.md Appendix A.2 / A.3.

Reference call sites: ``o3.wigner_3j`` (``e3_layers/nn/output.py:172`` and, implicitly, every
``o3.TensorProduct`` / ``o3.FullyConnectedTensorProduct`` built at ``nn/pointwise.py:78`` and
``nn/message_passing.py:83``); ``o3.SphericalHarmonics`` (``nn/embedding.py:163-165``).

e3nn 0.4.4 is not available here -> the tensors are re-derived analytically (the e3nn >= 0.5
on-the-fly recipe): SU(2) Clebsch-Gordan coefficients by the Racah formula, rotated into the
real basis by the change of basis Q_l, Frobenius-normalised.  Sign convention risk R1 (odd
l1+l2+l3) is handled by ``SIGN_PRESET``.
"""
import functools
import math
from fractions import Fraction

import numpy as np
import torch

# "analytic": signs as produced by the construction below (e3nn >= 0.5).
# "e3nn044":  flips the triples whose first non-zero element (flat order, l1<=l2<=l3
#             representative) is negative in the analytic construction.
SIGN_PRESET = "analytic"


def _f(n):
    return math.factorial(int(round(n)))


def _su2_cg_coeff(j1, m1, j2, m2, j3, m3):
    """<j1 m1 j2 m2 | j3 m3>, Racah formula, integer spins only."""
    if m3 != m1 + m2:
        return 0.0
    vmin = int(max(-j1 + j2 + m3, -j1 + m1, 0))
    vmax = int(min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3))
    C = math.sqrt(
        (2.0 * j3 + 1.0)
        * Fraction(
            _f(j3 + j1 - j2) * _f(j3 - j1 + j2) * _f(j1 + j2 - j3) * _f(j3 + m3) * _f(j3 - m3),
            _f(j1 + j2 + j3 + 1) * _f(j1 - m1) * _f(j1 + m1) * _f(j2 - m2) * _f(j2 + m2),
        )
    )
    S = 0
    for v in range(vmin, vmax + 1):
        S += (-1) ** int(v + j2 + m2) * Fraction(
            _f(j2 + j3 + m1 - v) * _f(j1 - m1 + v),
            _f(v) * _f(j3 - j1 + j2 - v) * _f(j3 + m3 - v) * _f(v + j1 - j2 - m3),
        )
    return C * float(S)


def _su2_cg(j1, j2, j3):
    mat = np.zeros((2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1), dtype=np.float64)
    if abs(j1 - j2) <= j3 <= j1 + j2:
        for m1 in range(-j1, j1 + 1):
            for m2 in range(-j2, j2 + 1):
                if abs(m1 + m2) <= j3:
                    mat[j1 + m1, j2 + m2, j3 + m1 + m2] = _su2_cg_coeff(j1, m1, j2, m2, j3, m1 + m2)
    return mat


def _real_to_complex(l):
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    s = 1 / math.sqrt(2)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = s
        q[l + m, l - abs(m)] = -1j * s
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m * s
        q[l + m, l - abs(m)] = 1j * (-1) ** m * s
    return (-1j) ** l * q


@functools.lru_cache(maxsize=None)
def _w3j_analytic(l1, l2, l3):
    Q1, Q2, Q3 = _real_to_complex(l1), _real_to_complex(l2), _real_to_complex(l3)
    C = _su2_cg(l1, l2, l3).astype(np.complex128)
    C = np.einsum("ij,kl,mn,ikn->jlm", Q1, Q2, np.conj(Q3.T), C)
    assert np.abs(C.imag).max() < 1e-10
    C = C.real
    n = np.linalg.norm(C)
    assert n > 0
    C = C / n
    C[np.abs(C) < 1e-14] = 0.0
    return C


def _sign_e3nn044(l1, l2, l3):
    """Global sign of the (l1,l2,l3) tensor under the 'e3nn044' preset relative to analytic."""
    ls = sorted([l1, l2, l3])
    C = _w3j_analytic(*ls).reshape(-1)
    centre = _w3j_analytic(*ls)[ls[0], ls[1], ls[2]]
    if abs(centre) > 1e-12:
        s_rep = 1.0 if centre > 0 else -1.0
    else:
        nz = C[np.abs(C) > 1e-12]
        s_rep = 1.0 if nz[0] > 0 else -1.0
    return s_rep


def wigner_3j_np(l1, l2, l3):
    """[2l1+1, 2l2+1, 2l3+1] float64 numpy array, Frobenius norm 1."""
    if not (abs(l1 - l2) <= l3 <= l1 + l2):
        raise ValueError("triangle inequality violated")
    C = _w3j_analytic(l1, l2, l3)
    if SIGN_PRESET == "e3nn044":
        C = C * _sign_e3nn044(l1, l2, l3)
    elif SIGN_PRESET != "analytic":
        raise ValueError(SIGN_PRESET)
    return C.copy()


def wigner_3j(l1, l2, l3, dtype=None, device=None):
    if dtype is None:
        dtype = torch.get_default_dtype()
    return torch.tensor(wigner_3j_np(l1, l2, l3), dtype=dtype, device=device)


# ------------------------------------------------------------------------------------------
# Real spherical harmonics, 'component' normalisation applied by the caller.
def spherical_harmonics_raw(l, x, y, z):
    """'integral/norm-free' polynomials of SURVEY A.2 *without* the sqrt(2l+1) factor,
    evaluated at (x, y, z) (tensors of equal shape).  Returns a list of 2l+1 tensors."""
    if l == 0:
        return [torch.ones_like(x)]
    if l == 1:
        return [x, y, z]
    s3 = math.sqrt(3.0)
    x2, y2, z2 = x * x, y * y, z * z
    x2z2 = x2 + z2
    sh20 = s3 * x * z
    sh21 = s3 * x * y
    sh22 = y2 - 0.5 * x2z2
    sh23 = s3 * y * z
    sh24 = (s3 / 2.0) * (z2 - x2)
    if l == 2:
        return [sh20, sh21, sh22, sh23, sh24]
    if l == 3:
        return [
            math.sqrt(5.0 / 6.0) * (sh20 * z + sh24 * x),
            math.sqrt(5.0) * sh20 * y,
            math.sqrt(3.0 / 8.0) * (4.0 * y2 - x2z2) * x,
            0.5 * y * (2.0 * y2 - 3.0 * x2z2),
            math.sqrt(3.0 / 8.0) * z * (4.0 * y2 - x2z2),
            math.sqrt(5.0) * sh24 * y,
            math.sqrt(5.0 / 6.0) * (sh24 * z - sh20 * x),
        ]
    raise NotImplementedError("oracle SH implemented for l <= 3")


def spherical_harmonics(ls, vec, normalize=True, normalization="component"):
    """e3nn ``o3.spherical_harmonics`` for a list of l's; vec [..., 3] -> [..., sum(2l+1)]."""
    if normalize:
        vec = torch.nn.functional.normalize(vec, dim=-1)  # v / max(|v|, 1e-12)
    x, y, z = vec[..., 0], vec[..., 1], vec[..., 2]
    out = []
    for l in ls:
        comps = spherical_harmonics_raw(l, x, y, z)
        if normalization == "component":
            c = math.sqrt(2 * l + 1)
        elif normalization == "norm":
            c = 1.0
        elif normalization == "integral":
            c = math.sqrt(2 * l + 1) / math.sqrt(4 * math.pi)
        else:
            raise ValueError(normalization)
        out += [c * t for t in comps]
    return torch.stack(out, dim=-1)


# ------------------------------------------------------------------------------------------
def wigner_D_from_matrix(l, p, R):
    """Real representation matrix of O(3) element R for irrep (l, p), in the basis of the
    SH / w3j above.  D_1 = R in (x, y, z) order; D_l by CG recursion D_{l-1} (x) D_1 -> D_l."""
    R = R.to(torch.float64)
    det = torch.linalg.det(R)
    Rp = R * torch.sign(det)  # proper rotation
    D = torch.ones(1, 1, dtype=torch.float64)
    for ll in range(1, l + 1):
        C = wigner_3j(ll - 1, 1, ll, dtype=torch.float64)
        D = (2 * ll + 1) * torch.einsum("ijk,abc,ia,jb->kc", C, C, D, Rp)
    if det < 0:
        D = D * p
    return D.to(R.dtype)


def rand_rotation(generator=None, dtype=torch.float64):
    """Random proper rotation matrix (QR of a Gaussian matrix, det fixed to +1)."""
    A = torch.randn(3, 3, generator=generator, dtype=torch.float64)
    Q, Rr = torch.linalg.qr(A)
    Q = Q * torch.sign(torch.diagonal(Rr))
    if torch.linalg.det(Q) < 0:
        Q[:, 0] = -Q[:, 0]
    return Q.to(dtype)
