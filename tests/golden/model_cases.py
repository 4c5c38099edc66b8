"""Inputs of the reference-code golden cases (tests/golden/make_golden_model.py writes the fixture from them with the reference's
own source; tests/test_oracle.py re-creates them).  Needs neither /root/reference nor a GPU."""
import numpy as np
import torch

from diffusion_edf_b200.synthetic import make_poses, make_scene, model_kwargs, model_kwargs_place

SAMPLE_KW = dict(diffusion_schedules=[[1.0, 0.15], [0.15, 0.09]], N_steps=[3, 3], timesteps=[0.04, 0.04], temperatures=[0.0, 0.0],
                 log_t_schedule=True, time_exponent_temp=1.0, time_exponent_alpha=0.5)


def weight_checksums(sd):
    keys = sorted(sd)
    tot = float(sum(v.double().abs().sum() for v in sd.values()))
    probe = [float(sd[k].double().sum()) for k in keys[:: max(1, len(keys) // 16)]]
    return np.array([len(keys), tot] + probe, dtype=np.float64)


def inputs(kind):
    if kind == "pick":
        x, rgb = make_scene(1500, seed=3, half_extent=12.0)
        Ts, t = make_poses(6, x, seed=3, spread=6.0)
        gx, gf = torch.zeros(8, 3), torch.zeros(8, 3)
    else:
        x, rgb = make_scene(1200, seed=5, half_extent=10.0)
        Ts, t = make_poses(4, x, seed=5, spread=5.0)
        gx, gf = make_scene(700, seed=6, half_extent=8.0)
        gx[:, 2] += 9.0                                      # part of the grasp cloud inside the keypoint bbox (z >= 8)
    return x, rgb, torch.zeros(len(x), dtype=torch.long), Ts, t, gx, gf, torch.zeros(len(gx), dtype=torch.long)


def seeded_oracle(kind):
    """The oracle model of a case: shipped kwargs, seed 0, and the parameters that are initialised to 0 / 1 (biases, layer-norm
    affine weights) randomised so that the golden numbers exercise them."""
    from oracle import model as OM
    torch.manual_seed(0)
    oracle = OM.MultiscaleScoreModel(**(model_kwargs() if kind == "pick" else model_kwargs_place()), deterministic=True).eval()
    with torch.no_grad():
        for n, p in oracle.named_parameters():
            if p.abs().sum() == 0:
                p.uniform_(-0.3, 0.3)
            elif n.endswith("affine_weight"):
                p.uniform_(0.7, 1.3)
    return oracle
