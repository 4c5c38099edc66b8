import sys, os, time, json
sys.path.insert(0, "/root/repo")
import torch
from diffusion_edf_b200 import FeaturedPoints, MultiscaleScoreModel, ops
from diffusion_edf_b200.synthetic import make_poses, make_scene, model_kwargs
ops.USE_TC_MLP = os.environ.get("TC", "1") == "1"
graph = os.environ.get("GRAPH", "1") == "1"
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MultiscaleScoreModel(**model_kwargs(), deterministic=True).eval().to(dev)
model.requires_grad_(False)
model.use_cuda_graph = graph
x, rgb = make_scene(10_000, seed=0)
T_seed, _ = make_poses(1024, x, seed=0)
key = FeaturedPoints(x.to(dev), rgb.to(dev), torch.zeros(len(x), dtype=torch.long, device=dev))
grasp = FeaturedPoints(torch.zeros(8, 3, device=dev), torch.zeros(8, 3, device=dev), torch.zeros(8, dtype=torch.long, device=dev))
kw = dict(diffusion_schedules=[[1.0, 0.15], [0.15, 0.09]], N_steps=[50, 50], timesteps=[0.04, 0.04], temperatures=[1.0, 1.0],
          log_t_schedule=True, time_exponent_temp=1.0, time_exponent_alpha=0.5)
with torch.no_grad():
    keys = model.get_key_pcd_multiscale(key); q = model.get_query_pcd(grasp)
    model.sample(T_seed.to(dev), keys, q, **{**kw, "N_steps": [3, 3]})
    torch.cuda.synchronize()
    if not graph: ops.PROFILE = {}
    t0 = time.perf_counter()
    traj = model.sample(T_seed.to(dev), keys, q, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("TC", ops.USE_TC_MLP, "graph", graph, "PDL", os.environ.get("DEDF_PDL", "1"), "ms/step", 1e3 * dt / 100)
    if not graph:
        prof, ops.PROFILE = ops.PROFILE, None
        print({k: round(sum(a.elapsed_time(b) for a, b in v) / 100, 4) for k, v in prof.items()})
