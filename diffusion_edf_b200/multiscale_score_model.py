"""MultiscaleScoreModel: drop-in for /root/reference/diffusion_edf/multiscale_score_model.py:25-117
(same constructor kwargs = ``model_kwargs`` of score_model_configs.yaml, same attributes)."""
from __future__ import annotations

from typing import Dict, List

import torch

from .gnn_data import FeaturedPoints
from .keypoint_extractor import KeypointExtractor, StaticKeypointModel
from .score_head import ScoreModelHead
from .score_model_base import ScoreModelBase
from .unet_feature_extractor import ForwardOnlyFeatureExtractor, UnetFeatureExtractor


class MultiscaleScoreModel(ScoreModelBase):
    def __init__(self, query_model: str, score_head_kwargs: Dict, key_kwargs: Dict, query_kwargs: Dict,
                 deterministic: bool = False):
        super().__init__()
        name = key_kwargs["feature_extractor_name"]
        if name == "UnetFeatureExtractor":
            self.key_model = UnetFeatureExtractor(**key_kwargs["feature_extractor_kwargs"], deterministic=deterministic)
        elif name == "ForwardOnlyFeatureExtractor":
            self.key_model = ForwardOnlyFeatureExtractor(**key_kwargs["feature_extractor_kwargs"], deterministic=deterministic)
        else:
            raise NotImplementedError(f"feature extractor {name!r}")
        if query_model == "StaticKeypointModel":
            self.query_model = StaticKeypointModel(**query_kwargs)
        elif query_model == "KeypointExtractor":
            self.query_model = KeypointExtractor(**query_kwargs, deterministic=deterministic)
        else:
            raise ValueError(f"Unknown query model: {query_model}")
        head_cls = ScoreModelHead
        if score_head_kwargs.get("ebm", False):       # critic configs (*_ebm): energy only, see score_head_ebm.py
            from .score_head_ebm import EbmScoreModelHead
            head_cls = EbmScoreModelHead
        kw = score_head_kwargs["key_tensor_field_kwargs"]
        # same in-place kwargs mutation as the reference (multiscale_score_model.py:79-85)
        assert "irreps_input" not in kw and "use_src_point_attn" not in kw and "use_dst_point_attn" not in kw
        kw["irreps_input"] = self.key_model.irreps_output
        kw["use_src_point_attn"] = False
        kw["use_dst_point_attn"] = False
        self.score_head = head_cls(max_time=float(score_head_kwargs["max_time"]),
                                         time_emb_mlp=score_head_kwargs["time_emb_mlp"], key_tensor_field_kwargs=kw,
                                         irreps_query_edf=self.query_model.irreps_output,
                                         lin_mult=float(score_head_kwargs["lin_mult"]), ang_mult=float(score_head_kwargs["ang_mult"]),
                                         edge_time_encoding=score_head_kwargs["edge_time_encoding"],
                                         query_time_encoding=score_head_kwargs["query_time_encoding"])
        self.lin_mult = self.score_head.lin_mult
        self.ang_mult = self.score_head.ang_mult

    def _key_pcd_multiscale(self, pcd: FeaturedPoints) -> List[FeaturedPoints]:
        return self.key_model(pcd)

    def _query_pcd(self, pcd: FeaturedPoints) -> FeaturedPoints:
        return self.query_model(pcd)
