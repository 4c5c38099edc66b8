"""Inputs of the reference-code golden cases (tests/golden/make_golden_model.py writes the fixture from them with the reference's
own source; tests/test_oracle.py re-creates them).  Needs neither /root/reference nor a GPU."""
import numpy as np
import torch

from diffusion_edf_b200.synthetic import make_poses, make_scene

SAMPLE_KW = dict(diffusion_schedules=[[1.0, 0.15], [0.15, 0.09]], N_steps=[3, 3], timesteps=[0.04, 0.04], temperatures=[0.0, 0.0],
                 log_t_schedule=True, time_exponent_temp=1.0, time_exponent_alpha=0.5)


def weight_checksums(sd):
    keys = sorted(sd)
    tot = float(sum(v.double().abs().sum() for v in sd.values()))
    probe = [float(sd[k].double().sum()) for k in keys[:: max(1, len(keys) // 16)]]
    return np.array([len(keys), tot] + probe, dtype=np.float64)


KINDS = ("pick", "place", "highres", "sapien_highres", "sapien_lowres", "ebm", "pick_c2", "pick_1024")
NO_LOSS = ("pick_1024",)        # cases without the get_train_loss record (1024 poses through the reference's source on the CPU: forward only)
# kind -> (kwargs function of diffusion_edf_b200.synthetic, model class name, has scores, has sample)
_SPEC = {"pick": ("model_kwargs", "MultiscaleScoreModel", True, True), "place": ("model_kwargs_place", "MultiscaleScoreModel", True, True),
         "highres": ("model_kwargs_highres", "MultiscaleScoreModel", True, False),
         "sapien_highres": ("model_kwargs_sapien_highres", "MultiscaleScoreModel", True, False),
         "sapien_lowres": ("model_kwargs_sapien_lowres", "PointAttentiveScoreModel", True, False),
         "ebm": ("model_kwargs_ebm", "MultiscaleScoreModel", False, False),
         # full-size cases on bench.py's scene: BASELINE config C2 (10k points, 128 poses) and the north-star width (1024 poses)
         "pick_c2": ("model_kwargs", "MultiscaleScoreModel", True, False), "pick_1024": ("model_kwargs", "MultiscaleScoreModel", True, False)}


def spec(kind):
    from diffusion_edf_b200 import synthetic
    fn, cls, has_scores, has_sample = _SPEC[kind]
    return getattr(synthetic, fn)(), cls, has_scores, has_sample


def inputs(kind):
    small_grasp = (torch.zeros(3, 3), torch.zeros(3, 3))
    if kind == "pick":
        x, rgb = make_scene(1500, seed=3, half_extent=12.0)
        Ts, t = make_poses(6, x, seed=3, spread=6.0)
        gx, gf = torch.zeros(8, 3), torch.zeros(8, 3)
    elif kind in ("pick_c2", "pick_1024"):
        x, rgb = make_scene(10_000, seed=0)
        Ts, t = make_poses(128 if kind == "pick_c2" else 1024, x, seed=0)
        gx, gf = torch.zeros(8, 3), torch.zeros(8, 3)
    elif kind == "place":
        x, rgb = make_scene(1200, seed=5, half_extent=10.0)
        Ts, t = make_poses(4, x, seed=5, spread=5.0)
        gx, gf = make_scene(700, seed=6, half_extent=8.0)
        gx[:, 2] += 9.0                                      # part of the grasp cloud inside the keypoint bbox (z >= 8)
    elif kind == "highres":
        x, rgb = make_scene(1500, seed=31, half_extent=12.0)
        Ts, t = make_poses(9, x, seed=31, spread=2.5)
        gx, gf = small_grasp
    elif kind == "sapien_highres":
        x, rgb = make_scene(1200, seed=41, half_extent=10.0)
        Ts, t = make_poses(7, x, seed=41, spread=2.5)
        gx, gf = small_grasp
    elif kind == "sapien_lowres":
        x, rgb = make_scene(1200, seed=51, half_extent=10.0)
        Ts, t = make_poses(5, x, seed=51, spread=3.0)
        gx, gf = small_grasp
    else:
        assert kind == "ebm"
        x, rgb = make_scene(1500, seed=21, half_extent=12.0)
        Ts, _ = make_poses(12, x, seed=21, spread=3.0)
        t = torch.ones(12)
        gx, gf = small_grasp
    return x, rgb, torch.zeros(len(x), dtype=torch.long), Ts, t, gx, gf, torch.zeros(len(gx), dtype=torch.long)


def seeded_oracle(kind):
    """The oracle model of a case: shipped kwargs, seed 0, and the parameters that are initialised to 0 / 1 (biases, layer-norm
    affine weights) randomised so that the golden numbers exercise them."""
    from oracle import model as OM
    kwargs, cls, _, _ = spec(kind)
    torch.manual_seed(0)
    oracle = getattr(OM, cls)(**kwargs, deterministic=True).eval()
    with torch.no_grad():
        for n, p in oracle.named_parameters():
            if p.abs().sum() == 0:
                p.uniform_(-0.3, 0.3)
            elif n.endswith("affine_weight"):
                p.uniform_(0.7, 1.3)
    return oracle


def feature_rows(n: int) -> slice:
    """Rows of a key scale's feature matrix kept in the fixture (all coordinates are kept): about 64 evenly spaced rows."""
    return slice(0, n, max(1, n // 64))


# BASELINE.json configs[0] ("C1"): MultiscaleTensorField, 1 layer, lmax = 1, 256-point cloud
C1_KWARGS = dict(irreps_input="16x0e+8x1e", irreps_output="16x0e+8x1e", irreps_sh="1x0e+1x1e", num_heads=4, fc_neurons=[-1, 16, 16],
                 length_emb_dim=16, irreps_query=None, edge_context_emb_dim=None, r_cluster_multiscale=[2.0, None],
                 length_enc_max_r=10.0, r_mincut_nonscalar_sh=0.1, n_layers=1)


def c1_inputs():
    g = torch.Generator().manual_seed(0)
    x0 = torch.rand(256, 3, generator=g) * 6 - 3
    f0 = torch.randn(256, 40, generator=g)
    xq = torch.rand(64, 3, generator=g) * 6 - 3
    return x0, f0, xq


def c1_seeded_oracle():
    from oracle import model as OM
    torch.manual_seed(0)
    tf = OM.MultiscaleTensorField(**C1_KWARGS).eval()
    with torch.no_grad():
        for n, p in tf.named_parameters():
            if p.abs().sum() == 0:
                p.uniform_(-0.3, 0.3)
            elif n.endswith("affine_weight"):
                p.uniform_(0.7, 1.3)
    return tf
