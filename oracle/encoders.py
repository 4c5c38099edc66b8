"""Length / time encoders, soft cutoffs and quaternion helpers restated.
Oracle = test infrastructure only.

Follows /root/reference/diffusion_edf/radial_func.py:10-17 (gaussian,
soft_step), :32-70 (soft_square_cutoff_2), :19-29 (soft_cutoff,
soft_square_cutoff), :168-227 (GaussianRadialBasis + _GaussianParamModule),
:231-278 (GaussianRadialBasisLayerFiniteCutoff), :291-316
(SinusoidalPositionEmbeddings); transforms.py:83-110 (quaternion_to_matrix),
:113-129, :132-144, :147-163, :198-209, :226-228, :271-307
(matrix_to_euler_angles, 'YXY' only).  Both reference files import in the
authoring container, so everything here is pinned by golden vectors produced
from the reference itself (tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn


# ---------------------------------------------------------------- cutoffs
def soft_step(x: torch.Tensor, n: int = 3) -> torch.Tensor:
    return (x > 0) * ((x < 1) * ((n + 1) * x.pow(n) - n * x.pow(n + 1)) + (x >= 1))


def soft_cutoff(x, thr: float = 0.8, n: int = 3):
    return 1 - soft_step((x - thr) / (1 - thr), n=n)


def soft_square_cutoff(x, thr: float = 0.8, n: int = 3, infinite: bool = False):
    if infinite:
        return soft_cutoff(x, thr=thr, n=n) * (x > 0.5) + soft_cutoff(1 - x, thr=thr, n=n) * (x <= 0.5)
    return (x > 0.5) + soft_cutoff(1 - x, thr=thr, n=n) * (x <= 0.5)


def soft_square_cutoff_2(x, ranges: Optional[Tuple[Optional[float], ...]], n: int = 3):
    if ranges is None:
        return x
    left_end, left_begin, right_begin, right_end = ranges
    div_l = 1.0 if left_end is None else left_begin - left_end
    div_r = 1.0 if right_end is None else right_end - right_begin
    if right_begin is not None and left_end is None:
        return 1 - soft_step((x - right_begin) / div_r, n=n)
    if left_end is not None and right_begin is None:
        return soft_step((x - left_end) / div_l, n=n)
    if right_begin is not None and left_end is not None:
        mid = 0.5 * (left_begin + right_begin)
        return (1 - soft_step((x - right_begin) / div_r, n=n)) * (x > mid) + soft_step((x - left_end) / div_l, n=n) * (x <= mid)
    return torch.ones_like(x)


# --------------------------------------------------------------- encoders
class _GaussianParamModule(nn.Module):
    def __init__(self, dim: int, max_weight: float):
        super().__init__()
        self.std_logit = nn.Parameter(torch.full((1, dim), math.log(math.exp(2.0 / dim) - 1), dtype=torch.float32))
        self.weight_logit = nn.Parameter(torch.full((1, dim), -math.log(max_weight / 1.0 - 1), dtype=torch.float32))
        self.mean = nn.Parameter(torch.linspace(0.0, 1.0, dim + 2, dtype=torch.float32)[1:-1].unsqueeze(0))
        self.weight_cap = max_weight * float(math.sqrt(dim))

    def forward(self):
        return self.mean + 0.0, F.softplus(self.std_logit) + 1e-5, torch.sigmoid(self.weight_logit) * self.weight_cap


class GaussianRadialBasis(nn.Module):
    def __init__(self, dim: int, max_val: float, min_val: float = 0.0):
        super().__init__()
        self.dim, self.max_val, self.min_val = int(dim), float(max_val), float(min_val)
        self.param_module = _GaussianParamModule(dim=dim, max_weight=4.0)

    def forward(self, dist):
        x = ((dist.unsqueeze(-1) - self.min_val) / (self.max_val - self.min_val)).expand(-1, self.dim)
        mean, std, weight = self.param_module()
        return torch.exp(-0.5 * (((x - mean) / std) ** 2)) * weight


class GaussianRadialBasisLayerFiniteCutoff(nn.Module):
    def __init__(self, num_basis: int, cutoff: float, soft_cutoff: bool = True, offset: Optional[float] = None,
                 cutoff_thr_ratio: float = 0.8, infinite: bool = False):
        super().__init__()
        self.num_basis, self.cutoff = num_basis, float(cutoff)
        self.offset = float(0.01 * self.cutoff if offset is None else offset)
        self.mean = nn.Parameter(torch.linspace(0, 1.0, num_basis + 2)[1:-1].unsqueeze(0))
        self.std_logit = nn.Parameter(torch.full((1, num_basis), math.log(math.exp(2.0 / num_basis) - 1)))
        self.max_weight = 4.0
        self.weight_logit = nn.Parameter(torch.full((1, num_basis), -math.log(self.max_weight / 1.0 - 1)))
        self.soft_cutoff, self.cutoff_thr_ratio = soft_cutoff, cutoff_thr_ratio
        self.normalizer, self.infinite = math.sqrt(num_basis), infinite

    def forward(self, dist):
        dist = ((dist - self.offset) / (self.cutoff - self.offset)).unsqueeze(-1)
        x = dist.expand(-1, self.num_basis)
        std = F.softplus(self.std_logit) + 1e-5
        x = torch.exp(-0.5 * (((x - self.mean) / std) ** 2))
        x = torch.sigmoid(self.weight_logit) * self.max_weight * x
        if self.soft_cutoff:
            x = x * soft_square_cutoff(dist, thr=self.cutoff_thr_ratio, infinite=self.infinite)
        return x * self.normalizer


class SinusoidalPositionEmbeddings(nn.Module):
    def __init__(self, dim: int, max_val: float, n: float = 10000.0):
        super().__init__()
        assert dim % 2 == 0
        self.dim, self.n, self.max_val = dim, float(n), float(max_val)

    def forward(self, x):
        x = x / self.max_val * self.n
        half = self.dim // 2
        k = math.log(self.n) / (half - 1)
        freq = torch.exp(torch.arange(half, dtype=x.dtype) * -k)
        e = x[..., None] * freq
        return torch.cat((e.sin(), e.cos()), dim=-1)


# ------------------------------------------------------------ quaternions
def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def quaternion_raw_multiply(a, b):
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    return torch.stack((aw * bw - ax * bx - ay * by - az * bz,
                        aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw), -1)


def quaternion_invert(q):
    return q * torch.tensor([1, -1, -1, -1], dtype=q.dtype)


def quaternion_apply(q, point):
    pq = torch.cat((point.new_zeros(point.shape[:-1] + (1,)), point), -1)
    return quaternion_raw_multiply(quaternion_raw_multiply(q, pq), quaternion_invert(q))[..., 1:]


def standardize_quaternion(q):
    return torch.where(q[..., 0:1] < 0, -q, q)


def normalize_quaternion(q):
    return q / torch.norm(q, dim=-1, keepdim=True)


def matrix_to_euler_yxy(m):
    """transforms.matrix_to_euler_angles(m, 'YXY') -> (…,3) = (alpha, beta, gamma)."""
    return torch.stack((torch.atan2(m[..., 0, 1], m[..., 2, 1]),
                        torch.acos(m[..., 1, 1]),
                        torch.atan2(m[..., 1, 0], -m[..., 1, 2])), -1)


def transform_points(points, Ts):
    """edf_interface/edf_interface/data/pcd_utils.py:55-81, un-batched pcd, Ts (nT,7) -> (nT, N, 3)."""
    q, t = Ts[..., :4], Ts[..., 4:]
    n = points.shape[-2]
    return quaternion_apply(q.unsqueeze(-2).expand(-1, n, -1), points.unsqueeze(-3).expand(len(Ts), -1, -1)) + t.unsqueeze(-2)
