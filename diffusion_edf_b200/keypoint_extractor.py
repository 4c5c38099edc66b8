"""Query-side models.  Mirrors /root/reference/diffusion_edf/keypoint_extractor.py:22-47 (``StaticKeypointModel``: fixed
coordinates, learned features and weights -- pick configs) and :50-197 (``KeypointExtractor``: own UNet + FPS + two
tensor fields + a weight head -- place configs)."""
from __future__ import annotations

import torch
from torch import nn

import copy
from typing import Dict, Optional

from . import ops
from .gnn_data import FeaturedPoints
from .irreps import Irreps


class StaticKeypointModel(nn.Module):
    graph_safe = True       # forward() has no data-dependent shapes beyond the batch layout

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


class KeypointExtractor(nn.Module):
    graph_safe = False      # the bbox filter selects a data-dependent subset: MultiscaleScoreModel.forward stays eager

    def __init__(self, feature_extractor_kwargs: Dict, tensor_field_kwargs: Dict, keypoint_kwargs: Dict,
                 feature_extractor_name: str = "UnetFeatureExtractor", weight_activation: str = "sigmoid",
                 weight_mult: Optional[float] = None, deterministic: bool = False):
        super().__init__()
        from .multiscale_tensor_field import MultiscaleTensorField
        from .unet_feature_extractor import UnetFeatureExtractor
        if feature_extractor_name != "UnetFeatureExtractor":
            raise NotImplementedError(f"feature extractor {feature_extractor_name!r}")
        if weight_activation not in ("sigmoid", "none"):
            raise NotImplementedError(f"weight_activation {weight_activation!r}")     # 'softmax' is dead code in the reference (:135-136)
        self.deterministic = deterministic
        self.pool_ratio = float(keypoint_kwargs["pool_ratio"])
        self.keypoint_bbox = keypoint_kwargs.get("bbox", None)
        self.weight_pre_emb_dim = int(keypoint_kwargs["weight_pre_emb_dim"])
        assert self.weight_pre_emb_dim > 0
        if weight_mult is None:
            self.weight_mult_logit = None
        else:
            self.weight_mult_logit = nn.Parameter(torch.log(torch.exp(torch.tensor(float(weight_mult))) - 1))
        self.feature_extractor = UnetFeatureExtractor(**feature_extractor_kwargs, deterministic=deterministic)
        # same in-place kwargs mutation as the reference (:103-116)
        assert "irreps_input" not in tensor_field_kwargs and "irreps_query" not in tensor_field_kwargs
        assert "edge_context_emb_dim" not in tensor_field_kwargs
        tensor_field_kwargs["irreps_input"] = feature_extractor_kwargs["irreps_output"]
        tensor_field_kwargs["irreps_query"] = None
        tensor_field_kwargs["edge_context_emb_dim"] = None
        self.tensor_field = MultiscaleTensorField(**tensor_field_kwargs)
        tensor_field_kwargs["irreps_output"] = f"{self.weight_pre_emb_dim}x0e"
        self.weight_field = MultiscaleTensorField(**tensor_field_kwargs)
        self.weight_post = nn.Sequential(nn.LayerNorm(self.weight_pre_emb_dim), nn.SiLU(inplace=True),
                                         nn.Linear(self.weight_pre_emb_dim, 1),
                                         nn.Sigmoid() if weight_activation == "sigmoid" else nn.Identity())
        self.use_sigmoid = weight_activation == "sigmoid"
        self.irreps_output = Irreps(self.tensor_field.irreps_output)

    def get_query_points(self, src_points: FeaturedPoints) -> FeaturedPoints:
        x, b = src_points.x, src_points.b
        if self.keypoint_bbox is not None:
            bb = torch.tensor(self.keypoint_bbox, dtype=x.dtype, device=x.device)
            idx = ((x >= bb[:, 0]) * (x <= bb[:, 1])).all(dim=-1).nonzero().squeeze(-1)     # data-dependent size: one host sync
            x, b = x.index_select(0, idx), b.index_select(0, idx)
        sel = ops.fps(x.contiguous(), b.contiguous(), self.pool_ratio, random_start=not self.deterministic)
        x, b = ops.gather_rows(x.contiguous(), sel), b.index_select(0, sel)
        return FeaturedPoints(x=x, f=torch.empty_like(x), b=b, w=None)

    def forward(self, input_points: FeaturedPoints, max_neighbors: int = 1000) -> FeaturedPoints:
        keys = self.feature_extractor(input_points)
        q = self.get_query_points(input_points)
        sources = self.tensor_field.encode_sources(keys)
        out = self.tensor_field(query_points=q, input_points_multiscale=keys, max_neighbors=max_neighbors, sources=sources)
        wsrc = self.weight_field.encode_sources(keys)
        wf = self.weight_field(query_points=q, input_points_multiscale=keys, max_neighbors=max_neighbors, sources=wsrc).f
        ln, lin = self.weight_post[0], self.weight_post[2]
        w = ops.weight_post(wf.contiguous(), ln.weight.detach(), ln.bias.detach(), lin.weight.detach().reshape(-1).contiguous(),
                            lin.bias.detach(), self.use_sigmoid,
                            None if self.weight_mult_logit is None else self.weight_mult_logit.detach().reshape(1))
        return FeaturedPoints(x=out.x, f=out.f, b=out.b, w=w)
