"""Public carrier types of the score network, identical to the reference's
(/root/reference/diffusion_edf/gnn_data.py:12-16 ``FeaturedPoints``, :117-124 ``GraphEdge``),
plus ``TransformPcd`` (:80-100) on the CUDA path.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch
from torch import nn

from . import ops
from .irreps import Irreps


class FeaturedPoints(NamedTuple):
    x: torch.Tensor                    # (N, 3) position
    f: torch.Tensor                    # (N, F) feature, e3nn mul_ir layout
    b: torch.Tensor                    # (N,)   batch index (int64)
    w: Optional[torch.Tensor] = None   # (N,)   optional scalar weight


class GraphEdge(NamedTuple):
    edge_src: torch.Tensor
    edge_dst: torch.Tensor
    edge_length: Optional[torch.Tensor] = None
    edge_attr: Optional[torch.Tensor] = None
    edge_scalars: Optional[torch.Tensor] = None
    edge_weights: Optional[torch.Tensor] = None
    edge_logits: Optional[torch.Tensor] = None


def set_featured_points_attribute(points: FeaturedPoints, x=None, f=None, b=None, w="") -> FeaturedPoints:
    return FeaturedPoints(x=points.x if x is None else x, f=points.f if f is None else f,
                          b=points.b if b is None else b, w=points.w if isinstance(w, str) else w)


def detach_featured_points(points: FeaturedPoints) -> FeaturedPoints:
    return FeaturedPoints(x=points.x.detach(), f=points.f.detach(), b=points.b.detach(),
                          w=points.w.detach() if isinstance(points.w, torch.Tensor) else points.w)


def flatten_featured_points(points: FeaturedPoints) -> FeaturedPoints:
    return FeaturedPoints(x=points.x.reshape(-1, 3), f=points.f.reshape(-1, points.f.shape[-1]), b=points.b.reshape(-1),
                          w=points.w.reshape(-1) if points.w is not None else None)


def cat_featured_points(fp1: FeaturedPoints, fp2: FeaturedPoints) -> FeaturedPoints:
    w = None if (fp1.w is None or fp2.w is None) else torch.cat([fp1.w, fp2.w], dim=0)
    return FeaturedPoints(x=torch.cat([fp1.x, fp2.x]), f=torch.cat([fp1.f, fp2.f]), b=torch.cat([fp1.b, fp2.b]), w=w)


class _SliceAndTransform(nn.Module):
    """Holds the ``J`` buffer of the reference's SliceAndTransform (wigner.py:203-230) for state_dict
    compatibility; the kernel builds D(q) from R(q) directly and never reads it."""

    def __init__(self, l: int):
        super().__init__()
        self.register_buffer("J", torch.zeros(2 * l + 1, 2 * l + 1))


class TransformFeatureQuaternion(nn.Module):
    def __init__(self, irreps):
        super().__init__()
        self.irreps = Irreps(irreps)
        self.transforms = nn.ModuleList([_SliceAndTransform(l) for l, m in enumerate(self.irreps.m) if m])


class TransformPcd(nn.Module):
    """x' = R(q) x + t, f' = D(q) f for every pose: (nQ,·) x (nT,7) -> (nT, nQ, ·)."""

    def __init__(self, irreps):
        super().__init__()
        self.transform_features = TransformFeatureQuaternion(irreps)

    def forward(self, pcd: FeaturedPoints, Ts: torch.Tensor) -> FeaturedPoints:
        assert Ts.ndim == 2 and Ts.shape[-1] == 7, f"{Ts.shape}"
        n_t, n_q = Ts.shape[0], pcd.x.shape[0]
        x, f = ops.query_transform(Ts.contiguous(), pcd.x.contiguous(), pcd.f.contiguous(), self.transform_features.irreps.m)
        w = pcd.w.expand(n_t, -1) if isinstance(pcd.w, torch.Tensor) else None
        return FeaturedPoints(x=x.view(n_t, n_q, 3), f=f.view(n_t, n_q, -1), b=pcd.b.expand(n_t, -1), w=w)
