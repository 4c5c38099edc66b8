"""Query-side model.  Mirrors /root/reference/diffusion_edf/keypoint_extractor.py:22-47
(``StaticKeypointModel``: fixed coordinates, learned features and weights).  It has no
arithmetic beyond a sigmoid over ``len(keypoint_coords)`` numbers, done once per grasp."""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from .gnn_data import FeaturedPoints
from .irreps import Irreps


class StaticKeypointModel(nn.Module):
    def __init__(self, keypoint_coords, irreps_output):
        super().__init__()
        kc = torch.as_tensor(keypoint_coords, dtype=torch.float32)
        assert kc.ndim == 2 and kc.shape[-1] == 3, f"{kc.shape}"
        self.irreps_output = Irreps(irreps_output)
        self.register_buffer("keypoint_coords", kc)
        self.keypoint_features = nn.Parameter(torch.randn(len(kc), self.irreps_output.dim))
        self.keypoint_weights = nn.Parameter(torch.randn(len(kc)))

    def forward(self, input_points: FeaturedPoints) -> FeaturedPoints:
        b = input_points.b
        assert b.ndim == 1
        bu = ops.plan_value(lambda: torch.unique(b))      # device->host sync: recorded once under a CUDA-graph plan
        n = len(bu)
        return FeaturedPoints(x=self.keypoint_coords.repeat(n, 1), f=self.keypoint_features.repeat(n, 1),
                              b=bu.repeat(len(self.keypoint_coords)), w=torch.sigmoid(self.keypoint_weights).repeat(n))
