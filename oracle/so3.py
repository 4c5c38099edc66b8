"""SO(3) constants and functions of e3nn==0.4.4 restated (SURVEY.md App. A.2,
A.3, A.6).  Oracle = test infrastructure only.  e3nn is NOT in /root/reference
(pinned in its setup.py:26); what follows is its published algorithm:

* ``wigner_3j``: SU(2) Clebsch-Gordan coefficients (Racah formula) conjugated
  with the real<->complex change of basis that carries the extra ``(-i)^l``
  factor, real part, Frobenius-normalised.
* real spherical harmonics in e3nn's axis convention (y is the polar axis, so
  that the l=1 harmonics are (x, y, z) and D^1(R) = R).
* ``wigner_D(l, alpha, beta, gamma) = Z(alpha) J Z(beta) J Z(gamma)`` with the
  YXY Euler angles; ``J = D^l(S)``, S the half-turn about (1,1,0)/sqrt(2)
  (it exchanges the x and y axes, so ``J Z(b) J`` is the rotation about x).
"""
from __future__ import annotations

import functools
import math
from fractions import Fraction

import torch


# --------------------------------------------------------------------------
# 3j symbols
# --------------------------------------------------------------------------
def _su2_cg_coeff(j1, m1, j2, m2, j3, m3) -> float:
    if m3 != m1 + m2:
        return 0.0
    vmin = int(max(-j1 + j2 + m3, -j1 + m1, 0))
    vmax = int(min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3))

    def f(n):
        return math.factorial(round(n))

    c = (
        (2.0 * j3 + 1.0)
        * Fraction(
            f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) * f(j3 + m3) * f(j3 - m3),
            f(j1 + j2 + j3 + 1) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2),
        )
    ) ** 0.5
    s = 0
    for v in range(vmin, vmax + 1):
        s += (-1) ** int(v + j2 + m2) * Fraction(
            f(j2 + j3 + m1 - v) * f(j1 - m1 + v),
            f(v) * f(j3 - j1 + j2 - v) * f(j3 + m3 - v) * f(v + j1 - j2 - m3),
        )
    return float(c * s)


def _su2_cg(j1: int, j2: int, j3: int) -> torch.Tensor:
    mat = torch.zeros(2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1, dtype=torch.float64)
    if abs(j1 - j2) <= j3 <= j1 + j2:
        for m1 in range(-j1, j1 + 1):
            for m2 in range(-j2, j2 + 1):
                if abs(m1 + m2) <= j3:
                    mat[j1 + m1, j2 + m2, j3 + m1 + m2] = _su2_cg_coeff(j1, m1, j2, m2, j3, m1 + m2)
    return mat


def _real_to_complex(l: int) -> torch.Tensor:
    q = torch.zeros(2 * l + 1, 2 * l + 1, dtype=torch.complex128)
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
def wigner_3j(l1: int, l2: int, l3: int) -> torch.Tensor:
    """Real-basis 3j tensor of shape (2l1+1, 2l2+1, 2l3+1), float64, norm 1."""
    q1, q2, q3 = _real_to_complex(l1), _real_to_complex(l2), _real_to_complex(l3)
    c = _su2_cg(l1, l2, l3).to(torch.complex128)
    c = torch.einsum("ij,kl,mn,ikn->jlm", q1, q2, torch.conj(q3.T), c)
    assert float(c.imag.abs().max()) < 1e-9
    c = c.real.contiguous()
    return c / c.norm()


# --------------------------------------------------------------------------
# spherical harmonics  (o3.SphericalHarmonics(normalize=True, 'component'))
# --------------------------------------------------------------------------
def spherical_harmonics(lmax: int, vec: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    """(…,3) -> (…, (lmax+1)^2), l = 0..lmax, 'component' normalisation.

    ``normalize=True`` uses ``F.normalize`` (zero vector -> zeros), as e3nn does.
    """
    assert lmax <= 2
    if normalize:
        vec = torch.nn.functional.normalize(vec, dim=-1)
    x, y, z = vec[..., 0], vec[..., 1], vec[..., 2]
    out = [torch.ones_like(x)]
    if lmax >= 1:
        s3 = math.sqrt(3.0)
        out += [s3 * x, s3 * y, s3 * z]
    if lmax >= 2:
        s5, s15 = math.sqrt(5.0), math.sqrt(15.0)
        out += [
            s15 * x * z,
            s15 * x * y,
            s5 * (y * y - 0.5 * (x * x + z * z)),
            s15 * y * z,
            0.5 * s15 * (z * z - x * x),
        ]
    return torch.stack(out, dim=-1)


# --------------------------------------------------------------------------
# Wigner D
# --------------------------------------------------------------------------
def _fib_sphere(n: int) -> torch.Tensor:
    i = torch.arange(n, dtype=torch.float64) + 0.5
    phi = torch.acos(1 - 2 * i / n)
    th = math.pi * (1 + 5 ** 0.5) * i
    return torch.stack([torch.cos(th) * torch.sin(phi), torch.sin(th) * torch.sin(phi), torch.cos(phi)], -1)


def wigner_D_from_matrix(l: int, R: torch.Tensor) -> torch.Tensor:
    """D^l(R) defined by Y_l(R x) = D^l(R) Y_l(x); R (…,3,3) -> (…,2l+1,2l+1).

    Exact least-squares over sample directions (the Y_l are linearly
    independent); computed in float64 and cast back.
    """
    if l == 0:
        return torch.ones(R.shape[:-2] + (1, 1), dtype=R.dtype)
    pts = _fib_sphere(64)
    sl = slice(l * l, (l + 1) * (l + 1))
    Y = spherical_harmonics(l, pts)[..., sl]                     # (P, d)
    Rp = torch.einsum("...ij,pj->...pi", R.to(torch.float64), pts)
    YR = spherical_harmonics(l, Rp)[..., sl]                     # (…, P, d)
    pinv = torch.linalg.pinv(Y)                                  # (d, P)
    D = torch.einsum("...pa,bp->...ab", YR, pinv)                # YR = Y D^T
    return D.to(R.dtype)


@functools.lru_cache(maxsize=None)
def J_matrix(l: int) -> torch.Tensor:
    """e3nn ``_Jd[l]`` (float64): D^l of the half-turn exchanging x and y."""
    S = torch.tensor([[0.0, 1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, -1.0]], dtype=torch.float64)
    J = wigner_D_from_matrix(l, S)
    J = torch.where(J.abs() < 1e-12, torch.zeros_like(J), J)
    return J


def z_rot_mat(angle: torch.Tensor, l: int) -> torch.Tensor:
    """wigner.py:21-42."""
    M = angle.new_zeros((len(angle), 2 * l + 1, 2 * l + 1))
    inds = torch.arange(0, 2 * l + 1)
    rev = torch.arange(2 * l, -1, -1)
    freq = torch.arange(l, -l - 1, -1, dtype=angle.dtype)
    M[:, inds, rev] = torch.sin(freq * angle[:, None])
    M[:, inds, inds] = torch.cos(freq * angle[:, None])
    return M


def wigner_D_euler(l: int, alpha, beta, gamma) -> torch.Tensor:
    """wigner.py:44-81 (Xa @ J @ Xb @ J @ Xc)."""
    J = J_matrix(l).to(device=alpha.device, dtype=alpha.dtype)
    return z_rot_mat(alpha, l) @ J @ z_rot_mat(beta, l) @ J @ z_rot_mat(gamma, l)
