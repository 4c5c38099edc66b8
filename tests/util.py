import torch


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (b = reference)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_close(a, b, tol=1e-4, what=""):
    e = rel_err(a, b)
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e}"
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
