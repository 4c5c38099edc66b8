"""PointAttentiveScoreModel: drop-in for /root/reference/diffusion_edf/point_attentive_score_model.py:20-99 (the sapien*/*_lowres
configs): the KEY side is a KeypointExtractor too (own UNet + FPS + tensor fields + weight head), the score head sees that single
"scale" and weights every edge by its source point's weight after the softmax (use_src_point_attn=True)."""
from __future__ import annotations

from typing import Dict, List

from .gnn_data import FeaturedPoints
from .keypoint_extractor import KeypointExtractor, StaticKeypointModel
from .score_head import ScoreModelHead
from .score_model_base import ScoreModelBase


class PointAttentiveScoreModel(ScoreModelBase):
    def __init__(self, query_model: str, score_head_kwargs: Dict, key_kwargs: Dict, query_kwargs: Dict, deterministic: bool = False):
        super().__init__()
        self.key_model = KeypointExtractor(**key_kwargs, deterministic=deterministic)
        if query_model == "KeypointExtractor":
            self.query_model = KeypointExtractor(**query_kwargs, deterministic=deterministic)
        elif query_model == "StaticKeypointModel":
            self.query_model = StaticKeypointModel(**query_kwargs)
        else:
            raise ValueError(f"Unknown query model: {query_model}")
        if "lin_mult" not in score_head_kwargs or "ang_mult" not in score_head_kwargs:
            raise NotImplementedError()
        kw = score_head_kwargs["key_tensor_field_kwargs"]
        # same in-place kwargs mutation as the reference (point_attentive_score_model.py:69-74)
        assert "irreps_input" not in kw and "use_src_point_attn" not in kw and "use_dst_point_attn" not in kw
        kw["irreps_input"] = self.key_model.irreps_output
        kw["use_src_point_attn"] = True
        kw["use_dst_point_attn"] = False
        self.score_head = ScoreModelHead(max_time=float(score_head_kwargs["max_time"]), time_emb_mlp=score_head_kwargs["time_emb_mlp"],
                                         key_tensor_field_kwargs=kw, irreps_query_edf=self.query_model.irreps_output,
                                         lin_mult=float(score_head_kwargs["lin_mult"]), ang_mult=float(score_head_kwargs["ang_mult"]),
                                         edge_time_encoding=score_head_kwargs["edge_time_encoding"],
                                         query_time_encoding=score_head_kwargs["query_time_encoding"])
        self.lin_mult = self.score_head.lin_mult
        self.ang_mult = self.score_head.ang_mult

    def _key_pcd_multiscale(self, pcd: FeaturedPoints) -> List[FeaturedPoints]:
        return [self.key_model(pcd)]

    def _query_pcd(self, pcd: FeaturedPoints) -> FeaturedPoints:
        return self.query_model(pcd)
