"""Synthetic inputs of the benchmark shapes (BASELINE.md section 3 / SURVEY.md 8d): a surface-like
scene cloud in centimetres, random poses, a config loader.  Data generation only (CPU torch / numpy);
nothing here is on the compute path."""
from __future__ import annotations

import copy
import math
import os
from typing import Dict, Optional, Tuple

import torch

# model_kwargs of configs/panda_mug/pick_lowres/score_model_configs.yaml of the reference (identical to
# panda_bottle/pick_lowres), restated so that tests / bench do not need /root/reference at run time.
PANDA_MUG_PICK_LOWRES: Dict = {
    "score_head_kwargs": {
        "max_time": 1.0, "time_emb_mlp": [256, 128, 64], "ang_mult": 2.5, "lin_mult": 15.0,
        "edge_time_encoding": True, "query_time_encoding": False,
        "key_tensor_field_kwargs": {
            "irreps_output": "64x0e+32x1e+16x2e", "irreps_sh": "1x0e+1x1e+1x2e", "num_heads": 4,
            "fc_neurons": [-1, 128, 64], "length_emb_dim": 64, "r_cluster_multiscale": [5.0, 10.0, 20.0, None],
            "n_layers": 1, "irreps_mlp_mid": 3, "cutoff_method": "edge_attn", "r_mincut_nonscalar_sh": 0.3,
            "length_enc_max_r": 100.0,
        },
    },
    "key_kwargs": {
        "feature_extractor_name": "UnetFeatureExtractor",
        "feature_extractor_kwargs": {
            "irreps_input": "3x0e", "irreps_output": "64x0e+32x1e+16x2e", "n_scales": 4,
            "irreps_emb": ["32x0e+16x1e+8x2e", "32x0e+16x1e+8x2e", "64x0e+32x1e+16x2e", "64x0e+32x1e+16x2e"],
            "irreps_edge_attr": ["1x0e+1x1e+1x2e"] * 4, "num_heads": [4, 4, 4, 4],
            "fc_neurons": [[32, 16, 16], [32, 16, 16], [64, 32, 32], [64, 32, 32]], "n_layers": [2, 2, 2, 2],
            "pool_ratio": [0.2, 0.2, 0.2, 0.2], "radius": [3.0, None, None, None], "irreps_mlp_mid": 3,
            "pool_method": "fps", "attn_type": "mlp", "alpha_drop": 0.1, "proj_drop": 0.1, "drop_path_rate": 0.0,
            "n_layers_midstream": 2,
        },
    },
    "query_model": "StaticKeypointModel",
    "query_kwargs": {"irreps_output": "64x0e+32x1e+16x2e", "keypoint_coords": [[0.5, 0.5, 10.5], [-0.5, -0.5, 10.5]]},
}


def model_kwargs() -> Dict:
    """A fresh deep copy (the constructors mutate the dicts in place, like the reference's)."""
    return copy.deepcopy(PANDA_MUG_PICK_LOWRES)


def model_kwargs_place() -> Dict:
    """model_kwargs of configs/panda_mug/place_lowres/score_model_configs.yaml: same score head and key encoder as
    pick_lowres, query model = KeypointExtractor (own UNet with pool ratio 0.25, FPS 0.1 inside a bounding box, two
    tensor fields with radii [5, 10, 20, 40] cm and no context embedding, sigmoid weight head)."""
    kw = copy.deepcopy(PANDA_MUG_PICK_LOWRES)
    fe = copy.deepcopy(PANDA_MUG_PICK_LOWRES["key_kwargs"]["feature_extractor_kwargs"])
    fe["pool_ratio"] = [0.25, 0.25, 0.25, 0.25]
    kw["query_model"] = "KeypointExtractor"
    kw["query_kwargs"] = {
        "weight_activation": "sigmoid", "weight_mult": None,
        "keypoint_kwargs": {"pool_ratio": 0.1, "weight_pre_emb_dim": 64, "bbox": [[-30.0, 30.0], [-30.0, 30.0], [8.0, 100.0]]},
        "feature_extractor_kwargs": fe,
        "tensor_field_kwargs": {"irreps_output": "64x0e+32x1e+16x2e", "irreps_sh": "1x0e+1x1e+1x2e", "num_heads": 4,
                                "fc_neurons": [-1, 32, 32], "length_emb_dim": 64, "r_cluster_multiscale": [5.0, 10.0, 20.0, 40.0],
                                "n_layers": 1, "irreps_mlp_mid": 3, "cutoff_method": "edge_attn"},
    }
    return kw


def model_kwargs_highres() -> Dict:
    """model_kwargs of configs/panda_mug/pick_highres/score_model_configs.yaml: pick_lowres with key-field radii
    [3.5, 5, 6.5, 8] cm (all finite, no length_enc_max_r) and UNet pool ratio 0.25."""
    kw = copy.deepcopy(PANDA_MUG_PICK_LOWRES)
    kw["score_head_kwargs"]["key_tensor_field_kwargs"]["r_cluster_multiscale"] = [3.5, 5.0, 6.5, 8.0]
    kw["score_head_kwargs"]["key_tensor_field_kwargs"].pop("length_enc_max_r", None)
    kw["key_kwargs"]["feature_extractor_kwargs"]["pool_ratio"] = [0.25, 0.25, 0.25, 0.25]
    return kw


def model_kwargs_sapien_highres() -> Dict:
    """model_kwargs of configs/sapien/pick_highres/score_model_configs.yaml: ForwardOnlyFeatureExtractor (one scale, 7 layers,
    64x0e+32x1e+16x2e), time MLP [512, 256, 128] (192 edge scalars), one key-field radius of 6 cm."""
    return {
        "score_head_kwargs": {
            "max_time": 1.0, "time_emb_mlp": [512, 256, 128], "ang_mult": 2.5, "lin_mult": 15.0,
            "edge_time_encoding": True, "query_time_encoding": False,
            "key_tensor_field_kwargs": {
                "irreps_output": "64x0e+32x1e+16x2e", "irreps_sh": "1x0e+1x1e+1x2e", "num_heads": 4,
                "fc_neurons": [-1, 128, 64], "length_emb_dim": 64, "r_cluster_multiscale": [6.0],
                "n_layers": 1, "irreps_mlp_mid": 3, "cutoff_method": "edge_attn", "r_mincut_nonscalar_sh": 0.1,
            },
        },
        "key_kwargs": {
            "feature_extractor_name": "ForwardOnlyFeatureExtractor",
            "feature_extractor_kwargs": {
                "irreps_input": "3x0e", "irreps_output": "64x0e+32x1e+16x2e", "n_scales": 1,
                "irreps_emb": ["64x0e+32x1e+16x2e"], "irreps_edge_attr": ["1x0e+1x1e+1x2e"], "num_heads": [4],
                "fc_neurons": [[64, 32, 32]], "n_layers": [7], "pool_ratio": [0.25], "radius": [3.0], "irreps_mlp_mid": 3,
                "pool_method": "fps", "attn_type": "mlp", "alpha_drop": 0.1, "proj_drop": 0.1, "drop_path_rate": 0.0,
                "n_layers_midstream": 2,
            },
        },
        "query_model": "StaticKeypointModel",
        "query_kwargs": {"irreps_output": "64x0e+32x1e+16x2e", "keypoint_coords": [[0.0, -4.5, 10.0], [0.0, 4.5, 10.0]]},
    }


def model_kwargs_sapien_lowres() -> Dict:
    """model_kwargs of configs/sapien/pick_lowres/score_model_configs.yaml (model_name PointAttentiveScoreModel): the key side is a
    KeypointExtractor (4-scale UNet at 64x0e+32x1e+16x2e, pool 0.25, keypoint pool ratio 0.05, no bbox), the score head has one
    all-pairs scale with source-point attention and 192 edge scalars."""
    fe = {
        "irreps_input": "3x0e", "irreps_output": "64x0e+32x1e+16x2e", "n_scales": 4,
        "irreps_emb": ["64x0e+32x1e+16x2e"] * 4, "irreps_edge_attr": ["1x0e+1x1e+1x2e"] * 4, "num_heads": [4, 4, 4, 4],
        "fc_neurons": [[64, 32, 32]] * 4, "n_layers": [2, 2, 2, 2], "pool_ratio": [0.25] * 4, "radius": [3.0, None, None, None],
        "irreps_mlp_mid": 3, "pool_method": "fps", "attn_type": "mlp", "alpha_drop": 0.1, "proj_drop": 0.1, "drop_path_rate": 0.0,
        "n_layers_midstream": 2,
    }
    return {
        "score_head_kwargs": {
            "max_time": 1.0, "time_emb_mlp": [512, 256, 128], "ang_mult": 2.5, "lin_mult": 15.0,
            "edge_time_encoding": True, "query_time_encoding": False,
            "key_tensor_field_kwargs": {
                "irreps_output": "64x0e+32x1e+16x2e", "irreps_sh": "1x0e+1x1e+1x2e", "num_heads": 4,
                "fc_neurons": [-1, 128, 64], "length_emb_dim": 64, "r_cluster_multiscale": [None],
                "n_layers": 1, "irreps_mlp_mid": 3, "cutoff_method": "edge_attn", "r_mincut_nonscalar_sh": 0.1,
                "length_enc_max_r": 100.0,
            },
        },
        "key_kwargs": {
            "weight_activation": "sigmoid", "weight_mult": None,
            "keypoint_kwargs": {"pool_ratio": 0.05, "weight_pre_emb_dim": 64},
            "feature_extractor_name": "UnetFeatureExtractor", "feature_extractor_kwargs": fe,
            "tensor_field_kwargs": {"irreps_output": "64x0e+32x1e+16x2e", "irreps_sh": "1x0e+1x1e+1x2e", "num_heads": 4,
                                    "fc_neurons": [-1, 32, 32], "length_emb_dim": 64, "r_cluster_multiscale": [5.0, 10.0, 20.0, 40.0],
                                    "n_layers": 1, "irreps_mlp_mid": 3, "cutoff_method": "edge_attn"},
        },
        "query_model": "StaticKeypointModel",
        "query_kwargs": {"irreps_output": "64x0e+32x1e+16x2e", "keypoint_coords": [[0.0, -4.5, 10.0], [0.0, 4.5, 10.0]]},
    }


def model_kwargs_ebm() -> Dict:
    """model_kwargs of configs/panda_mug/pick_ebm/score_model_configs.yaml (the critic): pick_lowres with ``ebm: True``, no
    time encoding, key-field radii [3.5, 5, 6.5, 8] cm (all finite, no length_enc_max_r) and UNet pool ratio 0.25."""
    kw = copy.deepcopy(PANDA_MUG_PICK_LOWRES)
    sh = kw["score_head_kwargs"]
    sh["ebm"] = True
    sh["edge_time_encoding"] = False
    sh["query_time_encoding"] = False
    sh["key_tensor_field_kwargs"]["r_cluster_multiscale"] = [3.5, 5.0, 6.5, 8.0]
    sh["key_tensor_field_kwargs"].pop("length_enc_max_r", None)
    kw["key_kwargs"]["feature_extractor_kwargs"]["pool_ratio"] = [0.25, 0.25, 0.25, 0.25]
    return kw


def make_scene(n_points: int = 10_000, seed: int = 0, half_extent: float = 30.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Surface-like cloud (cm): table plane z=0 over [-h,h]^2 plus spheres / cylinders / boxes of radius 3-8 cm,
    jitter sigma 0.3 cm, 1 cm voxel average, random-subsampled / padded to exactly ``n_points``.  -> (x (N,3), rgb (N,3))"""
    g = torch.Generator().manual_seed(seed)
    h = half_extent
    dens = max(1.0, 2.2 * n_points / (4 * h * h))                      # points per cm^2 before the voxel filter
    parts = []
    n_plane = int(dens * 4 * h * h)
    parts.append(torch.stack([(torch.rand(n_plane, generator=g) * 2 - 1) * h, (torch.rand(n_plane, generator=g) * 2 - 1) * h,
                              torch.zeros(n_plane)], -1))
    for k in range(8):
        c = torch.tensor([(torch.rand(1, generator=g).item() * 2 - 1) * (h - 8), (torch.rand(1, generator=g).item() * 2 - 1) * (h - 8), 0.0])
        r = 3.0 + 5.0 * torch.rand(1, generator=g).item()
        kind = k % 3
        if kind == 0:      # sphere resting on the table
            n = int(dens * 4 * math.pi * r * r)
            v = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
            parts.append(c + torch.tensor([0, 0, r]) + r * v)
        elif kind == 1:    # upright cylinder (side + top)
            hgt = 6.0 + 10.0 * torch.rand(1, generator=g).item()
            n = int(dens * 2 * math.pi * r * hgt)
            th = torch.rand(n, generator=g) * 2 * math.pi
            parts.append(c + torch.stack([r * th.cos(), r * th.sin(), torch.rand(n, generator=g) * hgt], -1))
            n = int(dens * math.pi * r * r)
            rr, th = r * torch.rand(n, generator=g).sqrt(), torch.rand(n, generator=g) * 2 * math.pi
            parts.append(c + torch.stack([rr * th.cos(), rr * th.sin(), torch.full((n,), hgt)], -1))
        else:              # box (top + 4 sides)
            hgt = 4.0 + 8.0 * torch.rand(1, generator=g).item()
            n = int(dens * 4 * r * r)
            parts.append(c + torch.stack([(torch.rand(n, generator=g) * 2 - 1) * r, (torch.rand(n, generator=g) * 2 - 1) * r,
                                          torch.full((n,), hgt)], -1))
            for ax in range(2):
                for sgn in (-1.0, 1.0):
                    n = int(dens * 2 * r * hgt)
                    u, z = (torch.rand(n, generator=g) * 2 - 1) * r, torch.rand(n, generator=g) * hgt
                    fixed = torch.full((n,), sgn * r)
                    parts.append(c + (torch.stack([fixed, u, z], -1) if ax == 0 else torch.stack([u, fixed, z], -1)))
    pts = torch.cat(parts, 0)
    pts = pts + 0.3 * torch.randn(pts.shape, generator=g)
    # 1 cm voxel average
    vox = torch.floor(pts).to(torch.long)
    vox = vox - vox.min(0).values
    dims = vox.max(0).values + 1
    key = (vox[:, 0] * dims[1] + vox[:, 1]) * dims[2] + vox[:, 2]
    uniq, inv = torch.unique(key, return_inverse=True)
    cnt = torch.zeros(len(uniq)).index_add_(0, inv, torch.ones(len(pts)))
    ctr = torch.zeros(len(uniq), 3).index_add_(0, inv, pts) / cnt[:, None]
    perm = torch.randperm(len(ctr), generator=g)
    if len(ctr) >= n_points:
        ctr = ctr[perm[:n_points]]
    else:                  # pad with jittered copies
        extra = ctr[perm[torch.randint(len(ctr), (n_points - len(ctr),), generator=g)]]
        ctr = torch.cat([ctr, extra + 0.2 * torch.randn(extra.shape, generator=g)], 0)
    rgb = torch.rand(n_points, 3, generator=g)
    return ctr.float().contiguous(), rgb.float().contiguous()


def make_poses(n_poses: int, scene_x: torch.Tensor, seed: int = 0, spread: float = 15.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> Ts (nT,7) [qw,qx,qy,qz,x,y,z] with unit standardized quaternions, positions ~ N(centroid + [0,0,10], spread^2);
    time (nT,) ~ U(0.01, 1]."""
    g = torch.Generator().manual_seed(seed + 1)
    q = torch.nn.functional.normalize(torch.randn(n_poses, 4, generator=g), dim=-1)
    q = torch.where(q[:, :1] < 0, -q, q)
    x = scene_x.mean(0) + torch.tensor([0.0, 0.0, 10.0]) + spread * torch.randn(n_poses, 3, generator=g)
    t = 0.01 + 0.99 * torch.rand(n_poses, generator=g)
    return torch.cat([q, x], -1).float().contiguous(), t.float().contiguous()
