"""Pre-processing in front of the score network, on the device.

Mirrors the tensor-level functions of the reference's request path (agent.py:122-127):
  edf_interface/edf_interface/data/pcd_utils.py:123-152   voxel_filter(points, features, voxel_size, coord_reduction)
  edf_interface/edf_interface/data/preprocess.py:52-66     rescale (points * factor; poses: translation * factor)
  diffusion_edf/gnn_data.py:77-78                          pcd_to_featured_points
The reference's PointCloud / SE3 container classes and the on-disk demo format are out of scope."""
from __future__ import annotations

from typing import Tuple

import torch

from . import ops
from ._lib import ptr, stream
from .gnn_data import FeaturedPoints


def voxel_filter(points: torch.Tensor, features: torch.Tensor, voxel_size: float, coord_reduction: str = "average") -> Tuple[torch.Tensor, torch.Tensor]:
    """Drop-in for the reference's voxel_filter on CUDA tensors: (N,3), (N,F) -> (M,3), (M,F), ordered by ravelled voxel index."""
    if coord_reduction not in ("average", "center"):
        raise ValueError(f"Unknown coordinate reduction method: {coord_reduction}")
    assert points.device == features.device, f"{points.device} != {features.device}"
    assert points.ndim == 2 and points.shape[1] == 3 and features.ndim == 2 and len(features) == len(points)
    points, features = points.contiguous().float(), features.contiguous().float()
    n, F = points.shape[0], features.shape[1]
    dev = points.device
    if n == 0:
        return points.clone(), features.clone()
    mm = torch.empty(2, 3, dtype=torch.float32, device=dev)
    ops._call("dedf_bbox", ptr(points), n, ptr(mm[0]), ptr(mm[1]), stream())
    mins, maxs = mm.tolist()                                   # host read: sizes the dense grid (the reference syncs here too)
    vs = torch.tensor(voxel_size, dtype=torch.float32)
    # shape = max voxel index + 1, computed exactly like the device does (fp32 subtract, divide, truncate)
    shape = [int(torch.trunc((torch.tensor(maxs[d], dtype=torch.float32) - torch.tensor(mins[d], dtype=torch.float32)) / vs).item()) + 1 for d in range(3)]
    S = shape[0] * shape[1] * shape[2]
    key = torch.empty(n, dtype=torch.int32, device=dev)
    cnt = torch.zeros(S, dtype=torch.int32, device=dev)
    off = torch.empty(S, dtype=torch.int32, device=dev)
    rank = torch.empty(S, dtype=torch.int32, device=dev)
    n_occ = torch.zeros(1, dtype=torch.int32, device=dev)
    ops._call("dedf_voxel_count", ptr(points), n, ptr(mm[0]), float(voxel_size), shape[0], shape[1], shape[2], ptr(key, torch.int32),
              ptr(cnt, torch.int32), ptr(off, torch.int32), ptr(rank, torch.int32), ptr(n_occ, torch.int32), stream())
    m = int(n_occ.item())                                      # host read: output size
    out_p = torch.empty(m, 3, dtype=torch.float32, device=dev)
    out_f = torch.empty(m, F, dtype=torch.float32, device=dev)
    cursor = torch.zeros(S, dtype=torch.int32, device=dev)
    sorted_idx = torch.empty(n, dtype=torch.int32, device=dev)
    ops._call("dedf_voxel_reduce", ptr(points), ptr(features), n, F, ptr(mm[0]), float(voxel_size), shape[0], shape[1], shape[2],
              ptr(key, torch.int32), ptr(cnt, torch.int32), ptr(off, torch.int32), ptr(rank, torch.int32), ptr(cursor, torch.int32),
              ptr(sorted_idx, torch.int32), 1 if coord_reduction == "center" else 0, ptr(out_p), ptr(out_f), stream())
    return out_p, out_f


def pcd_to_featured_points(points: torch.Tensor, colors: torch.Tensor, batch_idx: int = 0) -> FeaturedPoints:
    """gnn_data.py:77-78: f = colours, b = batch_idx for every point."""
    return FeaturedPoints(x=points, f=colors, b=torch.full((len(points),), batch_idx, dtype=torch.long, device=points.device), w=None)


def downsample_and_rescale(points: torch.Tensor, colors: torch.Tensor, voxel_size: float, rescale_factor: float,
                           coord_reduction: str = "average") -> FeaturedPoints:
    """The agent's proc_fn for a point cloud (agent.py:122-127 with configs/*/preprocess.yaml: downsample 1 cm -> rescale x100):
    metres in, centimetres out."""
    p, c = voxel_filter(points, colors, voxel_size, coord_reduction)
    p = ops.add_scale(p, p, 0.5 * float(rescale_factor))        # (p + p) * f/2 = p * f on the device
    return pcd_to_featured_points(p, c)
