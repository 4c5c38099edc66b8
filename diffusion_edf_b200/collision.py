"""Collision-aware pre-place trajectory optimisation on the device: the post-processing step that takes the
sampled place poses and backs the grasped object out of the scene along the gradient of a point-cloud repulsion energy.

Mirrors the tensor-level functions of the reference (same names, argument meaning, defaults and error behaviour):
  edf_interface/edf_interface/utils/collision_utils.py:18-34     _check_pcd_collision / check_pcd_collision
  edf_interface/edf_interface/utils/collision_utils.py:40-110    _pcd_energy
  edf_interface/edf_interface/utils/collision_utils.py:116-147   _se3_adjoint_lie_grad (fused into the step kernel)
  edf_interface/edf_interface/utils/collision_utils.py:150-243   _optimize_pcd_collision_once / _optimize_pcd_collision_trajectory
  edf_interface/edf_interface/utils/collision_utils.py:245-287   optimize_pcd_collision_trajectory (incl. the voxel down-sampling)
  edf_interface/edf_interface/utils/manipulation_utils.py:82-107 compute_pre_place_trajectories
The reference's PointCloud / SE3 container classes are out of scope: clouds are (N, 3) tensors (or any object with
a ``.points`` tensor), poses (nPose, 7) tensors (or any object with a ``.poses`` tensor); trajectories come back as tensors.
Every step is three kernels of csrc/collision.cu (energy + gradient, pose update); no CPU fallback.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import ops
from ._lib import ptr, stream

_METHODS = {"knn": 0, "radius": 1}


def convert_to_tensor(x) -> torch.Tensor:
    """collision_utils.py:9-16 for duck-typed containers."""
    if isinstance(x, torch.Tensor):
        return x
    if hasattr(x, "points"):
        return x.points
    if hasattr(x, "poses"):
        return x.poses
    raise TypeError(f"expected a tensor, a point cloud (.points) or poses (.poses), got {type(x)}")


def _cloud(y: torch.Tensor) -> Tuple[torch.Tensor, int, int, int]:
    """-> (contiguous fp32 y, n_pose (or 1 for a shared cloud), n_y, pose stride in floats)"""
    assert (y.ndim == 2 or y.ndim == 3) and y.shape[-1] == 3, f"{y.shape}"
    y = y.detach().to(torch.float32).contiguous()
    if y.ndim == 2:
        return y, 1, y.shape[0], 0
    return y, y.shape[0], y.shape[1], 3 * y.shape[1]


def _check_pcd_collision(x: torch.Tensor, y: torch.Tensor, r: float) -> torch.Tensor:
    """(nPose,) bool: some point of y[pose] has a scene point within r."""
    assert x.ndim == 2 and x.shape[-1] == 3, f"{x.shape}"
    x = x.detach().to(torch.float32).contiguous()
    y, n_pose, n_y, stride = _cloud(y)
    hit = torch.empty(n_pose, dtype=torch.int32, device=x.device)
    ops._call("dedf_collision_check", ptr(x), x.shape[0], ptr(y), stride, None, n_pose, n_y, float(r), ptr(hit, torch.int32), stream())
    return hit >= 1


def check_pcd_collision(x, y, r: float) -> torch.Tensor:
    return _check_pcd_collision(x=convert_to_tensor(x), y=convert_to_tensor(y), r=r)


def _energy(x: torch.Tensor, y: torch.Tensor, stride: int, Ts: Optional[torch.Tensor], n_pose: int, n_y: int, cutoff_r: float,
            max_num_neighbor: int, eps: float, compute_grad: bool, cluster_method: str):
    if cluster_method not in _METHODS:
        raise ValueError(f"Unknown cluster method '{cluster_method}'")
    energy = torch.empty(n_pose, dtype=torch.float32, device=x.device)
    grad = torch.empty(n_pose, 6, dtype=torch.float32, device=x.device) if compute_grad else None
    ops._call("dedf_collision_energy", ptr(x), x.shape[0], ptr(y), stride, ptr(Ts), n_pose, n_y, float(cutoff_r), int(max_num_neighbor),
              float(eps), _METHODS[cluster_method], ptr(energy), ptr(grad), stream())
    return energy, grad


def _pcd_energy(x: torch.Tensor, y: torch.Tensor, cutoff_r: float, max_num_neighbor: int = 100, eps: float = 0.001,
                compute_grad: bool = True, cluster_method: str = "knn") -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """(energy (nPose,), grad (nPose, 6) or None): y is already in the scene frame, (nPose, nY, 3) or (nY, 3)."""
    assert x.ndim == 2 and x.shape[-1] == 3, f"{x.shape}"
    x = x.detach().to(torch.float32).contiguous()
    y, n_pose, n_y, stride = _cloud(y)
    return _energy(x, y, stride, None, n_pose, n_y, cutoff_r, max_num_neighbor, eps, compute_grad, cluster_method)


def _optimize_pcd_collision_once(x: torch.Tensor, y: torch.Tensor, Ts: torch.Tensor, dt: float, cutoff_r: float,
                                 max_num_neighbors: int = 100, eps: float = 0.01, cluster_method: str = "knn"):
    """One gradient step of every pose -> (new poses (nPose, 7), energy (nPose,)).  y: (nPose, nY, 3) in the object frame."""
    assert x.ndim == 2 and x.shape[-1] == 3, f"{x.shape}"
    assert y.ndim == 3 and y.shape[-1] == 3, f"{y.shape}"
    assert Ts.ndim == 2 and Ts.shape[-1] == 7, f"{Ts.shape}"
    assert len(Ts) == len(y), f"{Ts.shape}, {y.shape}"
    x = x.detach().to(torch.float32).contiguous()
    Ts = Ts.detach().to(torch.float32).contiguous()
    # an expanded (stride-0) cloud is passed once, not materialised per pose
    if y.stride(0) == 0:
        yy, stride = y[0].detach().to(torch.float32).contiguous(), 0
    else:
        yy, stride = y.detach().to(torch.float32).contiguous(), 3 * y.shape[1]
    energy, grad = _energy(x, yy, stride, Ts, len(Ts), y.shape[1], cutoff_r, max_num_neighbors, eps, True, cluster_method)
    new_pose = torch.empty_like(Ts)
    ops._call("dedf_collision_step", ptr(Ts), ptr(grad), len(Ts), float(dt), float(cutoff_r), ptr(new_pose), stream())
    return new_pose, energy


def _optimize_pcd_collision_trajectory(x: torch.Tensor, y: torch.Tensor, Ts: torch.Tensor, n_steps: int, dt: float, cutoff_r: float,
                                       max_num_neighbors: int = 100, eps: float = 0.01, cluster_method: str = "knn",
                                       revert_order: bool = False) -> torch.Tensor:
    """(nPose, n_steps, 7): the initial poses followed by n_steps - 1 optimisation steps (reversed if ``revert_order``)."""
    assert n_steps >= 1
    assert Ts.ndim == 2 and Ts.shape[-1] == 7, f"{Ts.shape}"
    n_poses = len(Ts)
    assert x.ndim == 2 and x.shape[-1] == 3, f"{x.shape}"
    assert (y.ndim == 2 or y.ndim == 3) and y.shape[-1] == 3, f"{y.shape}"
    if y.ndim == 2:
        y = y.expand(n_poses, -1, 3)
    traj = [Ts.detach().to(torch.float32)]
    for _ in range(n_steps - 1):
        new_pose, _ = _optimize_pcd_collision_once(x=x, y=y, Ts=traj[-1], dt=dt, cutoff_r=cutoff_r, max_num_neighbors=max_num_neighbors,
                                                   eps=eps, cluster_method=cluster_method)
        traj.append(new_pose)
    out = torch.stack(traj, dim=0).movedim(0, -2)
    if revert_order:
        out = torch.flip(out, dims=(-2,))
    return out


def optimize_pcd_collision_trajectory(x, y, Ts, n_steps: int, dt: float, cutoff_r: float, max_num_neighbors: int = 100, eps: float = 0.01,
                                      cluster_method: str = "knn", revert_order: bool = False, voxel_size: Optional[float] = None,
                                      voxel_coord_reduction: Optional[str] = None) -> List[torch.Tensor]:
    """List of nPose trajectories, each (n_steps, 7).  ``voxel_size``: both clouds are voxel-filtered first (preprocess.downsample)."""
    x, y, Ts = convert_to_tensor(x), convert_to_tensor(y), convert_to_tensor(Ts)
    if voxel_size is not None:
        from .preprocess import voxel_filter
        red = voxel_coord_reduction or "average"
        assert y.ndim == 2, "voxel down-sampling needs a single (nY, 3) grasp cloud"
        x, _ = voxel_filter(x, torch.zeros(len(x), 1, device=x.device), voxel_size, red)
        y, _ = voxel_filter(y, torch.zeros(len(y), 1, device=y.device), voxel_size, red)
    traj = _optimize_pcd_collision_trajectory(x=x, y=y, Ts=Ts, n_steps=n_steps, dt=dt, cutoff_r=cutoff_r, max_num_neighbors=max_num_neighbors,
                                              eps=eps, cluster_method=cluster_method, revert_order=revert_order)
    return [t for t in traj]


def compute_pre_place_trajectories(place_poses, scene_pcd, grasp_pcd, n_steps: int, dt: float, cutoff_r: float, max_num_neighbors: int = 100,
                                   eps: float = 0.01, cluster_method: str = "knn", voxel_size: Optional[float] = None,
                                   voxel_coord_reduction: Optional[str] = None) -> List[torch.Tensor]:
    """manipulation_utils.py:82-107: trajectories that END at the place poses (order reverted)."""
    return optimize_pcd_collision_trajectory(x=scene_pcd, y=grasp_pcd, Ts=place_poses, n_steps=n_steps, dt=dt, cutoff_r=cutoff_r,
                                             max_num_neighbors=max_num_neighbors, eps=eps, cluster_method=cluster_method, revert_order=True,
                                             voxel_size=voxel_size, voxel_coord_reduction=voxel_coord_reduction)
