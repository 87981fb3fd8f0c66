"""Real Clebsch-Gordan (Wigner-3j) tensors in the e3nn real basis, numpy fp64.

Used at BUILD time by csrc/gen_tp.py (constants of the unrolled kernels and the dense tables
of the generic kernel) and at run time by the host-side dense contractions (self-connection).
Definition: the SO(3)-invariant tensor of (l1,l2,l3) in the real basis in which
Y_1 = (x, y, z) (polar axis y), Frobenius norm 1.  Built by contracting the complex SU(2)
coupling coefficients with the real->complex change of basis U_l.
"""
import functools
from math import factorial as fac, sqrt

import numpy as np

SIGN_PRESETS = ("analytic", "e3nn044")


def _cg_complex(j1, j2, j3):
    """<j1 m1; j2 m2 | j3 m3> as an array [2j1+1, 2j2+1, 2j3+1] (integer j only)."""
    out = np.zeros((2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1))
    pref = (2 * j3 + 1) * fac(j1 + j2 - j3) * fac(j1 - j2 + j3) * fac(-j1 + j2 + j3) / fac(j1 + j2 + j3 + 1)
    for m1 in range(-j1, j1 + 1):
        for m2 in range(-j2, j2 + 1):
            m3 = m1 + m2
            if abs(m3) > j3:
                continue
            norm = sqrt(pref * fac(j3 + m3) * fac(j3 - m3) * fac(j1 - m1) * fac(j1 + m1) * fac(j2 - m2) * fac(j2 + m2))
            s = 0.0
            for k in range(0, j1 + j2 - j3 + 1):
                d = [k, j1 + j2 - j3 - k, j1 - m1 - k, j2 + m2 - k, j3 - j2 + m1 + k, j3 - j1 - m2 + k]
                if min(d) < 0:
                    continue
                den = 1
                for t in d:
                    den *= fac(t)
                s += (-1) ** k / den
            out[j1 + m1, j2 + m2, j3 + m3] = norm * s
    return out


def _u_real_to_complex(l):
    U = np.zeros((2 * l + 1, 2 * l + 1), dtype=complex)
    r = 1 / sqrt(2)
    U[l, l] = 1
    for m in range(1, l + 1):
        U[l - m, l + m] = r
        U[l - m, l - m] = -1j * r
        U[l + m, l + m] = (-1) ** m * r
        U[l + m, l - m] = 1j * (-1) ** m * r
    return (-1j) ** l * U


@functools.lru_cache(maxsize=None)
def _w3j(l1, l2, l3):
    C = _cg_complex(l1, l2, l3).astype(complex)
    U1, U2, U3 = _u_real_to_complex(l1), _u_real_to_complex(l2), _u_real_to_complex(l3)
    R = np.einsum("ai,bj,ck,abc->ijk", U1, U2, np.conj(U3), C)
    assert np.abs(R.imag).max() < 1e-12
    R = R.real / np.linalg.norm(R.real)
    R[np.abs(R) < 1e-14] = 0.0
    return R


def _preset_sign(l1, l2, l3):
    a = sorted((l1, l2, l3))
    R = _w3j(*a)
    c = R[a[0], a[1], a[2]]
    if abs(c) > 1e-12:
        return 1.0 if c > 0 else -1.0
    flat = R.reshape(-1)
    return 1.0 if flat[np.abs(flat) > 1e-12][0] > 0 else -1.0


def w3j(l1, l2, l3, preset="analytic"):
    if not abs(l1 - l2) <= l3 <= l1 + l2:
        raise ValueError("triangle rule")
    R = _w3j(l1, l2, l3)
    if preset == "e3nn044":
        R = R * _preset_sign(l1, l2, l3)
    elif preset != "analytic":
        raise ValueError(preset)
    return R.copy()
