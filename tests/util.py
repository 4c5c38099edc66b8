import json
import os

import torch


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (b = reference)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def elem_err(a: torch.Tensor, b: torch.Tensor, floor: float = 1e-2) -> float:
    """Element-wise error max_i |a_i - b_i| / (|b_i| + floor * max|b|): unlike the max-norm ``rel_err`` a small component that is wrong by
    100 % shows up (it is measured against its own magnitude, down to ``floor`` of the tensor's largest entry)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    scale = b.abs() + floor * b.abs().max().clamp_min(1e-30)
    return float(((a - b).abs() / scale).max())


def _log(what: str, e: float, ee: float, tol: float) -> None:
    path = os.environ.get("DEDF_PARITY_LOG")
    if path:
        with open(path, "a") as fh:
            fh.write(json.dumps({"test": os.environ.get("PYTEST_CURRENT_TEST", ""), "what": what, "max_norm": e, "elementwise": ee, "tol": tol}) + "\n")


def assert_close(a, b, tol=1e-4, what="", elem_factor: float = 50.0):
    """Both bounds must hold: max-norm relative error <= tol (BASELINE.json's north_star: 1e-4 relative fp32) AND the element-wise
    error (see ``elem_err``) <= elem_factor * tol.  fp32 round-off is O(1e-6 max|b|) in absolute terms whatever the size of the
    component, i.e. up to 100x larger relative to the smallest components the 1 % floor admits -- hence the factor; a component with
    a wrong sign, a missing term or a swapped index is off by O(1) of its own magnitude and fails it by two orders of magnitude."""
    e, ee = rel_err(a, b), elem_err(a, b)
    _log(what, e, ee, tol)
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e}"
    assert ee <= elem_factor * tol, f"{what}: element-wise error {ee:.3e} > {elem_factor * tol:.1e} (max-norm {e:.3e})"
    return e


def random_graph(n_src, n_dst, avg_deg, gen):
    """Random CSR-by-dst graph with ascending sources; returns (row_ptr int32, edge_src int32, edge_dst int32)."""
    deg = torch.poisson(torch.full((n_dst,), float(avg_deg)), generator=gen).long().clamp(max=n_src)
    deg[0] = 0   # always exercise an isolated destination
    rows, srcs = [], []
    for d in range(n_dst):
        k = int(deg[d])
        s = torch.randperm(n_src, generator=gen)[:k].sort().values
        srcs.append(s)
        rows.append(torch.full((k,), d, dtype=torch.long))
    edge_src = torch.cat(srcs)
    edge_dst = torch.cat(rows)
    row_ptr = torch.zeros(n_dst + 1, dtype=torch.long)
    row_ptr[1:] = deg.cumsum(0)
    return row_ptr.int(), edge_src.int(), edge_dst.int()
