"""One (or more) eager forwards of the C2 workload for ncu captures:  python profiles/run_forward.py [n_forwards] [n_poses]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusion_edf_b200 import FeaturedPoints, MultiscaleScoreModel
from diffusion_edf_b200.synthetic import make_poses, make_scene, model_kwargs

n_fw = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_poses = int(sys.argv[2]) if len(sys.argv) > 2 else 128
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
    for i in range(n_fw):
        (ang, lin), _ = model(Ts.to(dev), t.to(dev), key, grasp)
        torch.cuda.synchronize()
print("ok", float(ang.abs().sum()))
