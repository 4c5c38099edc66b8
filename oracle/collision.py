"""CPU oracle of the collision-aware pre-place trajectory optimisation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/ as the checker of diffusion_edf_b200/collision.py, never by the
product path.

Restates /root/reference/edf_interface/edf_interface/utils/collision_utils.py in plain torch (autograd for the energy gradient,
exactly as the reference does) together with the helpers it calls:
  _check_pcd_collision             collision_utils.py:18-34
  _pcd_energy                      collision_utils.py:40-110   (energy of the L1 distances of the k nearest / in-radius scene points,
                                                                gradient w.r.t. an infinitesimal rotation / translation of every pose)
  _se3_adjoint_lie_grad            collision_utils.py:116-147
  _optimize_pcd_collision_once     collision_utils.py:150-196
  _optimize_pcd_collision_trajectory  collision_utils.py:198-243
  se3._exp_map / se3._multiply     edf_interface/data/se3.py:13-23, :47-55 (pytorch3d se3_exp_map / matrix_to_quaternion,
                                   edf_interface/data/transforms.py:23-80, :425-561)
  compute_pre_place_trajectories   edf_interface/utils/manipulation_utils.py:82-107
Third-party ops restated: torch_cluster.knn (the k nearest x of every y by squared Euclidean distance, all of them if there are
fewer than k), torch_cluster.radius (oracle/graph.py), torch_scatter.scatter_sum (index_add_).

Pinned by tests/golden/make_golden_collision.py: the reference's own function sources (AST-extracted, decorators stripped, the three
third-party ops above supplied by stand-ins) on seeded inputs -> tests/golden/collision_golden.npz, held by
tests/test_oracle.py::test_collision_oracle_matches_reference_code_golden.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import encoders as enc
from . import graph as G


def knn(x: torch.Tensor, y: torch.Tensor, k: int) -> torch.Tensor:
    """torch_cluster.knn(x, y, k): LongTensor[2, E] = (y index, x index), the min(k, len(x)) nearest x of every y."""
    d2 = ((y[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    kk = min(k, len(x))
    idx = d2.topk(kk, dim=1, largest=False).indices
    rows = torch.arange(len(y)).repeat_interleave(kk)
    return torch.stack([rows, idx.reshape(-1)], dim=0)


def check_pcd_collision(x: torch.Tensor, y: torch.Tensor, r: float) -> torch.Tensor:
    """collision_utils.py:18-34: (nPose,) bool, True where any transformed grasp point has a scene point within r."""
    if y.ndim == 2:
        y = y.unsqueeze(0)
    n_poses, n_y = y.shape[:2]
    e = G.radius(x, y.reshape(-1, 3), r)
    pose_idx = e[0] // n_y
    n_edges = torch.zeros(n_poses, dtype=torch.long).index_add_(0, pose_idx, torch.ones_like(pose_idx))
    return n_edges >= 1


def pcd_energy(x: torch.Tensor, y: torch.Tensor, cutoff_r: float, max_num_neighbor: int = 100, eps: float = 0.001,
               compute_grad: bool = True, cluster_method: str = "knn") -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """collision_utils.py:40-110 -> (energy (nPose,), grad (nPose, 6) = d energy / d (rot xyz, trans xyz) at zero)."""
    if y.ndim == 2:
        y = y.unsqueeze(0)
    n_poses, n_y = y.shape[:2]
    x, y = x.detach(), y.detach()
    rot_y = trans_y = None
    if compute_grad:
        trans_y = torch.zeros(n_poses, 3, dtype=x.dtype).requires_grad_(True)
        y = y + trans_y.unsqueeze(-2)
        rot_y = torch.zeros(n_poses, 3, dtype=x.dtype).requires_grad_(True)
        dR = rot_y.unsqueeze(-1) * torch.eye(3, dtype=x.dtype)
        for i in range(3):
            y = y + torch.cross(dR[:, i:i + 1, :].expand_as(y), y, dim=-1)
    y = y.reshape(-1, 3)
    if cluster_method == "radius":
        e = G.radius(x, y.detach(), cutoff_r, max_num_neighbors=max_num_neighbor)
    elif cluster_method == "knn":
        e = knn(x, y.detach(), max_num_neighbor)
    else:
        raise ValueError(f"Unknown cluster method '{cluster_method}'")
    ey, ex = e[0], e[1]
    if len(ey) == 0:
        return torch.zeros(n_poses, dtype=x.dtype), torch.zeros(n_poses, 6, dtype=x.dtype)
    pose_idx = ey // n_y
    r = torch.norm(x[ex] - y[ey], dim=-1, p=1)
    if cluster_method == "knn":
        keep = r <= cutoff_r
        r, pose_idx = r[keep], pose_idx[keep]
    energy = cutoff_r / (r + eps * cutoff_r)
    energy = torch.zeros(n_poses, dtype=x.dtype).index_add_(0, pose_idx, energy)
    if not compute_grad:
        return energy.detach(), None
    energy.sum().backward()
    return energy.detach(), torch.cat([rot_y.grad.detach(), trans_y.grad.detach()], dim=-1)


def se3_adjoint_lie_grad(Ts: torch.Tensor, grad: torch.Tensor) -> torch.Tensor:
    """collision_utils.py:116-147."""
    qinv = enc.quaternion_invert(Ts[..., :4])
    g_r = grad[..., :3] - torch.cross(Ts[..., 4:], grad[..., 3:], dim=-1)
    return torch.cat([enc.quaternion_apply(qinv, g_r), enc.quaternion_apply(qinv, grad[..., 3:])], dim=-1)


def _hat(v: torch.Tensor) -> torch.Tensor:
    h = torch.zeros(len(v), 3, 3, dtype=v.dtype)
    x, y, z = v.unbind(1)
    h[:, 0, 1], h[:, 0, 2], h[:, 1, 0], h[:, 1, 2], h[:, 2, 0], h[:, 2, 1] = -z, y, z, -x, -y, x
    return h


def matrix_to_quaternion(m: torch.Tensor) -> torch.Tensor:
    """transforms.py:23-80 (pytorch3d): the best-conditioned of the four candidates."""
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m.reshape(-1, 9), dim=-1)
    q_abs = torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1).clamp(min=0).sqrt()
    cand = torch.stack([
        torch.stack([q_abs[:, 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[:, 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[:, 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[:, 3] ** 2], dim=-1)], dim=-2)
    cand = cand / (2.0 * q_abs[..., None].clamp(min=0.1))
    return cand[torch.arange(len(cand)), q_abs.argmax(dim=-1)]


def exp_map(lie: torch.Tensor, eps: float = 1e-4) -> torch.Tensor:
    """se3._exp_map (se3.py:47-55): (Rx, Ry, Rz, Vx, Vy, Vz) -> (n, 7) pose, through pytorch3d's se3_exp_map (transforms.py:425-561:
    squared rotation norm clamped at eps) and matrix_to_quaternion."""
    log_rot, log_t = lie[..., :3], lie[..., 3:]
    ang = (log_rot * log_rot).sum(1).clamp(min=eps).sqrt()
    K = _hat(log_rot)
    K2 = torch.bmm(K, K)
    eye = torch.eye(3, dtype=lie.dtype)[None]
    R = (ang.sin() / ang)[:, None, None] * K + ((1.0 - ang.cos()) / (ang * ang))[:, None, None] * K2 + eye
    V = eye + K * ((1 - ang.cos()) / ang ** 2)[:, None, None] + K2 * ((ang - ang.sin()) / ang ** 3)[:, None, None]
    t = torch.bmm(V, log_t[:, :, None])[:, :, 0]
    return torch.cat([matrix_to_quaternion(R), t], dim=-1)


def multiply(T1: torch.Tensor, T2: torch.Tensor) -> torch.Tensor:
    """se3._multiply (se3.py:13-23)."""
    q, x = enc.normalize_quaternion(T2[..., :4]), T2[..., 4:]
    x = enc.quaternion_apply(T1[..., :4], x) + T1[..., 4:]
    q = enc.normalize_quaternion(enc.quaternion_raw_multiply(T1[..., :4], q))
    return torch.cat([q, x], dim=-1)


def transform_points_batched(y: torch.Tensor, Ts: torch.Tensor) -> torch.Tensor:
    """pcd_utils.transform_points(batched_pcd=True) (pcd_utils.py:55-81): (nPose, nY, 3) x (nPose, 7)."""
    return enc.quaternion_apply(Ts[:, None, :4], y) + Ts[:, None, 4:]


def optimize_once(x, y, Ts, dt: float, cutoff_r: float, max_num_neighbors: int = 100, eps: float = 0.01, cluster_method: str = "knn"):
    """collision_utils.py:150-196 -> (new poses (nPose, 7), energy (nPose,))."""
    Ty = transform_points_batched(y, Ts)
    energy, grad = pcd_energy(x, Ty, cutoff_r, max_num_neighbor=max_num_neighbors, eps=eps, cluster_method=cluster_method)
    grad = se3_adjoint_lie_grad(Ts, grad)
    grad = grad * torch.tensor([1.0, 1.0, 1.0, cutoff_r, cutoff_r, cutoff_r], dtype=grad.dtype)
    disp = -grad * dt * cutoff_r
    return multiply(Ts, exp_map(disp)), energy


def optimize_trajectory(x, y, Ts, n_steps: int, dt: float, cutoff_r: float, max_num_neighbors: int = 100, eps: float = 0.01,
                        cluster_method: str = "knn", revert_order: bool = False) -> torch.Tensor:
    """collision_utils.py:198-243 -> (nPose, n_steps, 7)."""
    assert n_steps >= 1
    if y.ndim == 2:
        y = y.expand(len(Ts), -1, 3)
    traj = [Ts]
    for _ in range(n_steps - 1):
        new_pose, _ = optimize_once(x, y, traj[-1], dt, cutoff_r, max_num_neighbors, eps, cluster_method)
        traj.append(new_pose)
    out = torch.stack(traj, dim=0).movedim(0, -2)
    return torch.flip(out, dims=(-2,)) if revert_order else out
