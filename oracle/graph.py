"""torch_cluster / torch_scatter semantics restated.
Oracle = test infrastructure only.  The wheels are not under /root/reference
(un-pinned PyG wheels for torch 1.13.1+cu117, README.md:30); call sites:
graph_parser.py:339, connectivity.py:22,42,62, graph_attention.py:254-265.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch


def _sq_dist(y: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """Squared distances computed as sum of squared differences in the input
    dtype (the form torch_cluster's kernels use), shape (len(y), len(x))."""
    dx = x[None, :, 0] - y[:, None, 0]
    dy = x[None, :, 1] - y[:, None, 1]
    dz = x[None, :, 2] - y[:, None, 2]
    return (dx * dx + dy * dy) + dz * dz   # fixed association: the CUDA kernel uses the same, unfused


def radius(x: torch.Tensor, y: torch.Tensor, r: float, batch_x: Optional[torch.Tensor] = None,
           batch_y: Optional[torch.Tensor] = None, max_num_neighbors: int = 32) -> torch.Tensor:
    """For every y_j all x_i of the same batch with |x_i - y_j|^2 < r^2, ascending
    i, at most ``max_num_neighbors``; returns LongTensor[2, E] = (j, i) sorted by j."""
    rows, cols = [], []
    r2 = torch.tensor(float(r), dtype=x.dtype) ** 2
    chunk = max(1, int(4_000_000 // max(1, len(x))))
    for s in range(0, len(y), chunk):
        d2 = _sq_dist(y[s:s + chunk], x)
        mask = d2 < r2
        if batch_x is not None and batch_y is not None:
            mask &= batch_y[s:s + chunk, None] == batch_x[None, :]
        rank = mask.cumsum(dim=1)
        mask &= rank <= max_num_neighbors
        j, i = mask.nonzero(as_tuple=True)
        rows.append(j + s)
        cols.append(i)
    if not rows:
        return torch.zeros(2, 0, dtype=torch.long)
    return torch.stack([torch.cat(rows), torch.cat(cols)], dim=0)


def radius_graph(x: torch.Tensor, r: float, batch: Optional[torch.Tensor] = None, loop: bool = False,
                 max_num_neighbors: int = 32) -> torch.Tensor:
    """torch_cluster.radius_graph: radius(x, x) (one extra neighbour allowed when
    loops are then removed, as torch_cluster does), self pairs dropped."""
    e = radius(x, x, r, batch, batch, max_num_neighbors if loop else max_num_neighbors + 1)
    if not loop:
        keep = e[0] != e[1]
        e = e[:, keep]
    return e


def fps(src: torch.Tensor, batch: Optional[torch.Tensor], ratio: float, random_start: bool = False) -> torch.Tensor:
    """Farthest point sampling, deterministic start (index 0 of every batch
    segment), M = ceil(ratio * N) per segment, greedy arg-max of the running
    min squared distance; ties -> lowest index.  Returns indices in selection order."""
    assert not random_start, "oracle fps is the deterministic variant"
    if batch is None:
        batch = torch.zeros(len(src), dtype=torch.long)
    out = []
    for b in torch.unique(batch):
        idx = (batch == b).nonzero().squeeze(-1)
        pts = src[idx]
        n = len(pts)
        m = int(math.ceil(ratio * n))
        dist = torch.full((n,), float("inf"), dtype=src.dtype)
        cur = 0
        sel = []
        for _ in range(m):
            sel.append(cur)
            dx, dy, dz = pts[:, 0] - pts[cur, 0], pts[:, 1] - pts[cur, 1], pts[:, 2] - pts[cur, 2]
            d = (dx * dx + dy * dy) + dz * dz          # fixed association, unfused (matches the kernel)
            dist = torch.minimum(dist, d)
            cur = int(torch.argmax(dist))
        out.append(idx[torch.tensor(sel, dtype=torch.long)])
    return torch.cat(out)


def scatter_sum(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    out = src.new_zeros((dim_size,) + src.shape[1:])
    return out.index_add_(0, index, src)


def scatter_logsumexp(src: torch.Tensor, index: torch.Tensor, dim_size: int, eps: float = 1e-12) -> torch.Tensor:
    """torch_scatter.scatter_logsumexp over dim 0: max-shifted, eps inside the
    log, empty rows -> 0."""
    idx = index.reshape((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    mx = src.new_full((dim_size,) + src.shape[1:], float("-inf"))
    mx = mx.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
    mx_safe = torch.where(torch.isinf(mx), torch.zeros_like(mx), mx)
    s = scatter_sum(torch.exp(src - mx_safe[index]), index, dim_size)
    out = torch.log(s + eps) + mx_safe
    return torch.where(torch.isinf(mx), torch.zeros_like(out), out)


def voxel_filter(points: torch.Tensor, features: torch.Tensor, voxel_size: float, coord_reduction: str = "average"):
    """/root/reference/edf_interface/edf_interface/data/pcd_utils.py:123-152 restated without numpy / torch_scatter:
    torch_scatter.scatter(sum) over the ravelled voxel index == index_add_ (sequential on the CPU: ascending point order),
    np.ravel_multi_index == C-order ravel, ``nonzero()`` == ascending ravelled index."""
    mins = points.min(dim=-2).values
    vox_idx = torch.div((points - mins), voxel_size, rounding_mode="trunc").type(torch.long)
    shape = vox_idx.max(dim=-2).values + 1
    raveled = (vox_idx[:, 0] * shape[1] + vox_idx[:, 1]) * shape[2] + vox_idx[:, 2]
    size = int(shape[0] * shape[1] * shape[2])
    n_pts = torch.zeros(size, dtype=torch.long).index_add_(0, raveled, torch.ones_like(raveled))
    nonzero = n_pts.nonzero().squeeze(-1)
    n_pts = n_pts[nonzero]
    feat = torch.zeros(size, features.shape[1], dtype=features.dtype).index_add_(0, raveled, features)[nonzero]
    feat = feat / n_pts.unsqueeze(-1)
    if coord_reduction == "center":
        iz = nonzero % shape[2]
        iy = (nonzero // shape[2]) % shape[1]
        ix = nonzero // (shape[2] * shape[1])
        coord = torch.stack([ix, iy, iz], dim=-1)
        coord = coord * voxel_size + mins + (voxel_size / 2)
    elif coord_reduction == "average":
        coord = torch.zeros(size, 3, dtype=points.dtype).index_add_(0, raveled, points)[nonzero]
        coord = coord / n_pts.unsqueeze(-1)
    else:
        raise ValueError(f"Unknown coordinate reduction method: {coord_reduction}")
    return coord, feat
