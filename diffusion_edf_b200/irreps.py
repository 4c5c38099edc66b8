"""Irreps bookkeeping for the product path.

Every irreps the reference's configs use on the hot path is an even-parity triple
``m0 x0e + m1 x1e + m2 x2e`` with l ascending (features 64x0e+32x1e+16x2e /
32x0e+16x1e+8x2e, inputs 3x0e, spherical harmonics 1x0e+1x1e+1x2e and the irreps
derived from them).  The fused kernels are specialised for that family; other even-parity l<=2
irreps (BASELINE config C1: 16x0e+8x1e with the l<=1 harmonics) run the tensor field through the un-fused
kernels with a table-driven depthwise tensor product (``dtp_paths``).
"""
from __future__ import annotations

from typing import Tuple, Union


class Irreps:
    """An even-parity, l<=2, l-ascending irreps triple (m0, m1, m2)."""

    __slots__ = ("m",)

    def __init__(self, spec: Union[str, "Irreps", Tuple[int, int, int], None]):
        if isinstance(spec, Irreps):
            self.m = spec.m
            return
        if hasattr(spec, "m") and isinstance(getattr(spec, "m"), tuple):
            self.m = tuple(spec.m)
            return
        if isinstance(spec, (tuple, list)) and len(spec) == 3 and all(isinstance(v, int) for v in spec):
            self.m = (int(spec[0]), int(spec[1]), int(spec[2]))
            return
        m = [0, 0, 0]
        last_l = -1
        s = str(spec).replace(" ", "")
        if s:
            for tok in s.split("+"):
                mul, ir = tok.split("x") if "x" in tok else ("1", tok)
                l, p = int(ir[:-1]), ir[-1]
                if p != "e" or l > 2:
                    raise NotImplementedError(f"only even-parity l<=2 irreps are supported on the CUDA path, got {spec!r}")
                if l < last_l:
                    raise NotImplementedError(f"irreps must be sorted by l, got {spec!r}")
                last_l = l
                m[l] += int(mul)
        self.m = (m[0], m[1], m[2])

    @property
    def dim(self) -> int:
        return self.m[0] + 3 * self.m[1] + 5 * self.m[2]

    @property
    def num_irreps(self) -> int:
        return sum(self.m)

    @property
    def lmax(self) -> int:
        return 2 if self.m[2] else (1 if self.m[1] else 0)

    def __eq__(self, other) -> bool:
        return isinstance(other, Irreps) and self.m == other.m

    def __hash__(self):
        return hash(self.m)

    def __mul__(self, k: int) -> "Irreps":      # (irreps * k) sorted and simplified
        return Irreps((self.m[0] * k, self.m[1] * k, self.m[2] * k))

    def div(self, k: int) -> "Irreps":
        if any(v % k for v in self.m):
            raise ValueError(f"{self} cannot be divided by {k}")
        return Irreps((self.m[0] // k, self.m[1] // k, self.m[2] // k))

    def __repr__(self) -> str:
        return "+".join(f"{v}x{l}e" for l, v in enumerate(self.m) if v) or ""

    def __str__(self) -> str:
        return self.__repr__()


def dtp_paths(irr: Irreps, sh_lmax: int = 2, lo_filter: Tuple[int, ...] = (0, 1, 2)):
    """Paths of DepthwiseTensorProduct(irr, sh l <= sh_lmax, filter) in CREATION order (tensor_product_rescale.py:352-382):
    for every input irrep l1 (with m[l1] > 0), every harmonic l2, every |l1 - l2| <= lo <= l1 + l2 that is in the output filter
    (0e is always kept).  -> [(l1, l2, lo, mul, w_off, ch_off)]: ``w_off`` = offset of the path's ``mul`` weights, ``ch_off`` = the
    path's first channel inside the lo block of the sorted, simplified output."""
    raw = []
    for l1 in range(3):
        if not irr.m[l1]:
            continue
        for l2 in range(sh_lmax + 1):
            for lo in range(abs(l1 - l2), l1 + l2 + 1):
                if lo <= 2 and (lo in lo_filter or lo == 0):
                    raw.append((l1, l2, lo, irr.m[l1]))
    out, w_off, ch = [], 0, [0, 0, 0]
    for l1, l2, lo, mul in raw:
        out.append((l1, l2, lo, mul, w_off, ch[lo]))
        w_off += mul
        ch[lo] += mul
    return out


def _lo_filter(irr_out) -> Tuple[int, ...]:
    return (0, 1, 2) if irr_out is None else tuple(l for l in range(3) if Irreps(irr_out).m[l])


def dtp_out(irr: Irreps, sh_lmax: int = 2, irr_out=None) -> Irreps:
    """Output irreps (sorted, simplified) of the depthwise TP of ``irr`` with the l <= sh_lmax harmonics, restricted to the l's of
    ``irr_out`` (None: all l <= 2 kept -- SURVEY.md App. E.1: (m0+m1+m2, m0+3m1+2m2, m0+2m1+3m2) for the full case)."""
    m = [0, 0, 0]
    for _, _, lo, mul, _, _ in dtp_paths(irr, sh_lmax, _lo_filter(irr_out)):
        m[lo] += mul
    return Irreps(tuple(m))


def dtp_numel(irr: Irreps, sh_lmax: int = 2, irr_out=None) -> int:
    """Weights of the depthwise TP (one per path and input channel): 3 m0 + 6 m1 + 6 m2 in the full case."""
    return sum(p[3] for p in dtp_paths(irr, sh_lmax, _lo_filter(irr_out)))


def is_fused_family(irr: Irreps) -> bool:
    """2G x0e + G x1e + G/2 x2e with G in (16, 32): the family the fused edge kernels are specialised for (every shipped config)."""
    m0, m1, m2 = irr.m
    return m1 in (16, 32) and m0 == 2 * m1 and m1 == 2 * m2


def gate_pre(irr: Irreps) -> Irreps:
    """Input irreps of Gate for output ``irr``: scalars + one gate per non-scalar irrep."""
    m0, m1, m2 = irr.m
    return Irreps((m0 + m1 + m2, m1, m2))
