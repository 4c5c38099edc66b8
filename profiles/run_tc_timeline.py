import os, sys, ctypes
sys.path.insert(0, "/root/repo")
import torch
from diffusion_edf_b200 import FeaturedPoints, MultiscaleScoreModel, ops, _lib
from diffusion_edf_b200.synthetic import make_poses, make_scene, model_kwargs
n_poses = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MultiscaleScoreModel(**model_kwargs(), deterministic=True).eval().to(dev)
model.requires_grad_(False); model.use_cuda_graph = False
x, rgb = make_scene(10_000, seed=0)
Ts, t = make_poses(n_poses, x, seed=0)
key = FeaturedPoints(x.to(dev), rgb.to(dev), torch.zeros(len(x), dtype=torch.long, device=dev))
grasp = FeaturedPoints(torch.zeros(8, 3, device=dev), torch.zeros(8, 3, device=dev), torch.zeros(8, dtype=torch.long, device=dev))
lib = _lib.load()
lib.dedf_tc_set_debug.argtypes = [ctypes.c_void_p]
with torch.no_grad():
    keys = model.get_key_pcd_multiscale(key); q = model.get_query_pcd(grasp)
    src = model.score_head.key_tensor_field.encode_sources(keys)
    Tsd, td = Ts.to(dev), t.to(dev)
    for _ in range(2):
        model.score_head(Ts=Tsd, key_pcd_multiscale=keys, query_pcd=q, time=td[:1], sources=src, shared_time=True)
    dbg = torch.zeros(192, dtype=torch.int64, device=dev)
    lib.dedf_tc_set_debug(dbg.data_ptr())
    model.score_head(Ts=Tsd, key_pcd_multiscale=keys, query_pcd=q, time=td[:1], sources=src, shared_time=True)
    torch.cuda.synchronize()
    lib.dedf_tc_set_debug(None)
d = dbg.cpu().tolist()
for name, off in (("epi(t0)", 0), ("producer(t128)", 64), ("mma(t160)", 128)):
    v = [z for z in d[off:off+64] if z]
    print(name, [z - v[0] for z in v])
