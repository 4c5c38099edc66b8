"""Per-entry-point CUDA-event breakdown of ONE score-head step (the inner loop of sample()) at n_poses poses on the C2 scene.
    python profiles/run_head_breakdown.py [n_poses=1024]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusion_edf_b200 import FeaturedPoints, MultiscaleScoreModel, ops
from diffusion_edf_b200.synthetic import make_poses, make_scene, model_kwargs

n_poses = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MultiscaleScoreModel(**model_kwargs(), deterministic=True).eval().to(dev)
model.requires_grad_(False)
model.use_cuda_graph = False
x, rgb = make_scene(10_000, seed=0)
Ts, t = make_poses(n_poses, x, seed=0)
key = FeaturedPoints(x.to(dev), rgb.to(dev), torch.zeros(len(x), dtype=torch.long, device=dev))
grasp = FeaturedPoints(torch.zeros(8, 3, device=dev), torch.zeros(8, 3, device=dev), torch.zeros(8, dtype=torch.long, device=dev))
with torch.no_grad():
    keys = model.get_key_pcd_multiscale(key)
    q = model.get_query_pcd(grasp)
    src = model.score_head.key_tensor_field.encode_sources(keys)
    Tsd, td = Ts.to(dev), t.to(dev)
    for _ in range(3):
        model.score_head(Ts=Tsd, key_pcd_multiscale=keys, query_pcd=q, time=td[:1], sources=src, shared_time=True)
    torch.cuda.synchronize()
    g = ops.radius_csr(src[0], ops.query_transform(Tsd, q.x.contiguous(), q.f.contiguous(), (64, 32, 16))[0], [5.0, 10.0, 20.0, None], src_off=src[2])
    ops.PROFILE = {}
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        model.score_head(Ts=Tsd, key_pcd_multiscale=keys, query_pcd=q, time=td[:1], sources=src, shared_time=True)
    e1.record()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
out = {k: round(sum(a.elapsed_time(b) for a, b in v) / n, 4) for k, v in prof.items()}
print(json.dumps({"n_poses": n_poses, "edges": g.n_edges, "edges_per_query_node": g.n_edges / (n_poses * 2), "eager_ms_per_step": e0.elapsed_time(e1) / n,
                  "ms_per_entry_point": dict(sorted(out.items(), key=lambda kv: -kv[1]))}))
