"""First GPU call of the next round: A/B of the experimental kernels that were written after round 1's GPU minutes ran out
(DESIGN 12, item 0).  One JSON line:

    python profiles/run_probe.py            # ~20 s on one B200

* value_reduce_split_kernel (DEDF_VR_SPLIT=1): max relative difference against the default kernel on the real head graph
  (1024 poses, 86 k edges) and on a UNet level, then the per-entry-point breakdown of a score-head step at 128 / 1024 poses
  and the C2 forward (CUDA-graph replay) with and without the flag.
Nothing here is a benchmark value; bench.py is."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusion_edf_b200 import FeaturedPoints, MultiscaleScoreModel, ops
from diffusion_edf_b200.synthetic import make_poses, make_scene, model_kwargs

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MultiscaleScoreModel(**model_kwargs(), deterministic=True).eval().to(dev)
model.requires_grad_(False)
x, rgb = make_scene(10_000, seed=0)
key = FeaturedPoints(x.to(dev), rgb.to(dev), torch.zeros(len(x), dtype=torch.long, device=dev))
grasp = FeaturedPoints(torch.zeros(8, 3, device=dev), torch.zeros(8, 3, device=dev), torch.zeros(8, dtype=torch.long, device=dev))
out = {}


def flag(on: bool):
    if on:
        os.environ["DEDF_VR_SPLIT"] = "1"
    else:
        os.environ.pop("DEDF_VR_SPLIT", None)


def head_step(n_poses: int, split: bool):
    """eager per-entry-point times of one score-head step (same recipe as run_head_breakdown.py)"""
    flag(split)
    Ts, t = make_poses(n_poses, x, seed=0)
    Tsd, td = Ts.to(dev), t.to(dev)
    model.use_cuda_graph = False
    with torch.no_grad():
        keys = model.get_key_pcd_multiscale(key)
        q = model.get_query_pcd(grasp)
        src = model.score_head.key_tensor_field.encode_sources(keys)
        for _ in range(3):
            res = model.score_head(Ts=Tsd, key_pcd_multiscale=keys, query_pcd=q, time=td[:1], sources=src, shared_time=True)
        torch.cuda.synchronize()
        ops.PROFILE = {}
        n = 10
        for _ in range(n):
            res = model.score_head(Ts=Tsd, key_pcd_multiscale=keys, query_pcd=q, time=td[:1], sources=src, shared_time=True)
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
    model.use_cuda_graph = True
    flag(False)
    ms = {k: round(sum(a.elapsed_time(b) for a, b in v) / n, 4) for k, v in prof.items()}
    return ms, [r.clone() for r in res]


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


for n_poses in (128, 1024):
    ms0, r0 = head_step(n_poses, False)
    ms1, r1 = head_step(n_poses, True)
    out[f"head_{n_poses}"] = {"value_reduce_us": [1e3 * ms0["dedf_value_reduce"], 1e3 * ms1["dedf_value_reduce"]],
                              "step_sum_us": [1e3 * sum(ms0.values()), 1e3 * sum(ms1.values())],
                              "rel_diff_ang": rel(r1[0], r0[0]), "rel_diff_lin": rel(r1[1], r0[1])}

# C2 forward, CUDA-graph replay, L2 not flushed (A/B only)
Ts, t = make_poses(128, x, seed=0)
Tsd, td = Ts.to(dev), t.to(dev)
res = {}
for split in (False, True):
    flag(split)
    model._graphs.clear()                                         # force a re-capture so that the flag takes effect
    with torch.no_grad():
        for _ in range(4):
            (ang, lin), _ = model(Tsd, td, key, grasp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            (ang, lin), _ = model(Tsd, td, key, grasp)
        e1.record()
        torch.cuda.synchronize()
    res[split] = (e0.elapsed_time(e1) / 20, ang.clone(), lin.clone())
flag(False)
out["c2_forward_ms"] = [res[False][0], res[True][0]]
out["c2_rel_diff"] = [rel(res[True][1], res[False][1]), rel(res[True][2], res[False][2])]
out["note"] = "pairs are [default, DEDF_VR_SPLIT=1]; the flag is read at launch time, a captured CUDA graph keeps the kernel it was captured with"
print(json.dumps(out))
