"""Minimal irreps bookkeeping (restates the subset of e3nn==0.4.4 ``o3.Irreps``
the reference uses).  Oracle = test infrastructure only.

An irreps object is an ordered list of entries ``(mul, l, p)``; a feature vector
is the concatenation of the entries, each laid out ``[mul][2l+1]`` row-major.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple, Union

Entry = Tuple[int, int, int]  # (mul, l, p) with p = +1 ('e') or -1 ('o')


class Irreps:
    def __init__(self, spec: Union[str, "Irreps", Sequence[Entry], None] = None):
        if spec is None:
            self.entries: List[Entry] = []
        elif isinstance(spec, Irreps):
            self.entries = list(spec.entries)
        elif isinstance(spec, str):
            self.entries = []
            s = spec.replace(" ", "")
            if s:
                for tok in s.split("+"):
                    if "x" in tok:
                        mul, ir = tok.split("x")
                    else:
                        mul, ir = "1", tok
                    self.entries.append((int(mul), int(ir[:-1]), 1 if ir[-1] == "e" else -1))
        else:
            self.entries = [(int(m), int(l), int(p)) for (m, l, p) in spec]

    # -- basic protocol ---------------------------------------------------
    def __iter__(self):
        return iter(self.entries)

    def __len__(self):
        return len(self.entries)

    def __getitem__(self, i):
        return self.entries[i]

    def __eq__(self, other):
        return isinstance(other, Irreps) and self.entries == other.entries

    def __hash__(self):
        return hash(tuple(self.entries))

    def __add__(self, other: "Irreps") -> "Irreps":
        return Irreps(self.entries + Irreps(other).entries)

    def __mul__(self, n: int) -> "Irreps":
        # e3nn: ``Irreps * int`` repeats the list (graph_attention.py:165, gnn_block.py:106)
        return Irreps(self.entries * int(n))

    def __repr__(self):
        return "+".join(f"{m}x{l}{'e' if p == 1 else 'o'}" for m, l, p in self.entries)

    # -- derived quantities -----------------------------------------------
    @property
    def dim(self) -> int:
        return sum(m * (2 * l + 1) for m, l, _ in self.entries)

    @property
    def num_irreps(self) -> int:
        return sum(m for m, _, _ in self.entries)

    @property
    def lmax(self) -> int:
        return max(l for _, l, _ in self.entries)

    def slices(self) -> List[slice]:
        out, i = [], 0
        for m, l, _ in self.entries:
            out.append(slice(i, i + m * (2 * l + 1)))
            i += m * (2 * l + 1)
        return out

    def simplify(self) -> "Irreps":
        out: List[Entry] = []
        for m, l, p in self.entries:
            if m == 0:
                continue
            if out and out[-1][1] == l and out[-1][2] == p:
                out[-1] = (out[-1][0] + m, l, p)
            else:
                out.append((m, l, p))
        return Irreps(out)

    def count(self, l: int, p: int = 1) -> int:
        return sum(m for m, ll, pp in self.entries if ll == l and pp == p)


def sort_even_first(irreps: Irreps):
    """equiformer/tensor_product_rescale.py:385-392 -- stable sort by (l, -p).

    Returns (sorted irreps, p, inv) with ``p[i]`` = new position of old entry i.
    """
    keyed = sorted((l, -p, i, m) for i, (m, l, p) in enumerate(irreps))
    inv = tuple(i for _, _, i, _ in keyed)
    perm = [0] * len(inv)
    for new, old in enumerate(inv):
        perm[old] = new
    return Irreps([(m, l, -np_) for l, np_, _, m in keyed]), tuple(perm), inv


def multiply_irreps(irreps: Irreps, mult: float) -> Irreps:
    """irreps_utils.py:7-17 (strict)."""
    out = []
    for m, l, p in irreps:
        if round(m * mult) != m * mult:
            raise ValueError(f"{irreps} cannot be multiplied by {mult}")
        out.append((round(m * mult), l, p))
    return Irreps(out)


def selection_rule(l1: int, p1: int, l2: int, p2: int) -> Iterable[Tuple[int, int]]:
    """Irrep product ``ir1 * ir2`` in ascending l (e3nn ``Irrep.__mul__``)."""
    for l in range(abs(l1 - l2), l1 + l2 + 1):
        yield (l, p1 * p2)


def irreps2gate(irreps: Irreps):
    """equiformer/tensor_product_rescale.py:188-238."""
    scal = Irreps([(m, l, p) for m, l, p in irreps if l == 0 and p == 1]).simplify()
    gated = Irreps([(m, l, p) for m, l, p in irreps if not (l == 0 and p == 1)]).simplify()
    gates = Irreps([(m, 0, 1) for m, _, _ in gated]).simplify()
    return scal, gates, gated
