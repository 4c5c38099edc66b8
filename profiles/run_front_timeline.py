"""Phase timeline of dedf_head_front (every CTA's %globaltimer stamps) inside a replayed denoise step.
    python profiles/run_front_timeline.py [n_poses=128]"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusion_edf_b200 import FeaturedPoints, MultiscaleScoreModel, _lib
from diffusion_edf_b200.denoise import DenoiseGraph
from diffusion_edf_b200.synthetic import make_poses, make_scene, model_kwargs

n_poses = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MultiscaleScoreModel(**model_kwargs(), deterministic=True).eval().to(dev)
model.requires_grad_(False)
x, rgb = make_scene(10_000, seed=0)
T_seed, _ = make_poses(n_poses, x, seed=0)
key = FeaturedPoints(x.to(dev), rgb.to(dev), torch.zeros(len(x), dtype=torch.long, device=dev))
grasp = FeaturedPoints(torch.zeros(8, 3, device=dev), torch.zeros(8, 3, device=dev), torch.zeros(8, dtype=torch.long, device=dev))
lib = _lib.load()
lib.dedf_head_front_set_debug.argtypes = [ctypes.c_void_p]
dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
with torch.no_grad():
    keys = model.get_key_pcd_multiscale(key)
    q = model.get_query_pcd(grasp)
    src = model.score_head.key_tensor_field.encode_sources(keys)
    n_steps = 64
    rows = [[0.5, 1e-3, 1e-3, 1.0]] * n_steps
    rows_all = model.score_head.time_rows_for(torch.full((n_steps,), 0.5, device=dev))
    lib.dedf_head_front_set_debug(dbg.data_ptr())
    DenoiseGraph.STEPS_PER_GRAPH = 1
    dg = DenoiseGraph(model, n_poses, n_steps, src, q, False, dev)
    dg.run(T_seed.double().to(dev), src, q, rows, rows_all, None, 0)
    torch.cuda.synchronize()
    lib.dedf_head_front_set_debug(None)
d = dbg.cpu().view(148, 8)
d = d[d[:, 0] > 0]
t0 = int(d[:, 0].min())
names = ["start", "staged", "dep_wait_over", "points", "counted", "barrier", "csr", "filled"]
rel = (d - t0).double() / 1e3
print(json.dumps({"n_poses": n_poses, "ctas": int(d.shape[0]),
                  "us_since_first_cta_start": {n: {"min": round(float(rel[:, i].min()), 2), "median": round(float(rel[:, i].median()), 2),
                                                   "max": round(float(rel[:, i].max()), 2)} for i, n in enumerate(names)}}))
