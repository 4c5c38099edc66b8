"""diffusion_edf_b200: B200-native (sm_100a) implementation of Diffusion-EDF's SE(3)-equivariant
score network (MultiscaleScoreModel.forward / ScoreModelBase.sample) behind the reference's own
module API.  Host code is PyTorch (device memory, streams, torch.distributed); the arithmetic is
hand-written CUDA in libdedf.so, bound through the C ABI of include/dedf.h.  No CPU fallback."""
from .gnn_data import FeaturedPoints, GraphEdge, TransformPcd  # noqa: F401
from .multiscale_score_model import MultiscaleScoreModel  # noqa: F401
from .point_attentive_score_model import PointAttentiveScoreModel  # noqa: F401
from .score_head_ebm import EbmScoreModelHead  # noqa: F401
from .score_head import ScoreModelHead  # noqa: F401
from .multiscale_tensor_field import MultiscaleTensorField  # noqa: F401
from .unet_feature_extractor import ForwardOnlyFeatureExtractor, UnetFeatureExtractor  # noqa: F401
from .keypoint_extractor import KeypointExtractor, StaticKeypointModel  # noqa: F401

__all__ = ["FeaturedPoints", "GraphEdge", "TransformPcd", "MultiscaleScoreModel", "PointAttentiveScoreModel", "ScoreModelHead",
           "EbmScoreModelHead", "MultiscaleTensorField", "UnetFeatureExtractor", "ForwardOnlyFeatureExtractor", "StaticKeypointModel",
           "KeypointExtractor"]
