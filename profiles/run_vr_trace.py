"""Phase timeline of dedf_value_reduce inside a replayed denoise step (clock64 stamps of thread 0, CTA 0):
    python profiles/run_vr_trace.py build      # here: profiles/_ab/libdedf_trace.so = the library with -DDEDF_VR_TRACE
    DEDF_LIB=profiles/_ab/libdedf_trace.so python profiles/run_vr_trace.py [n_poses=128]      # on the GPU box
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "build":
    import __graft_entry__ as g
    csrc = os.path.join(ROOT, "diffusion_edf_b200", "csrc")
    os.makedirs(os.path.join(ROOT, "profiles", "_ab"), exist_ok=True)
    g.build()
    obj = os.path.join(ROOT, "profiles", "_ab", "edge_trace.o")
    subprocess.check_call([g.NVCC, "-O3", "-std=c++17", "-lineinfo", *g.ARCH, "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
                           "-DDEDF_VR_TRACE", "-c", os.path.join(csrc, "edge.cu"), "-o", obj])
    objs = [os.path.join(csrc, s.replace(".cu", ".o")) for s in g.SOURCES if s != "edge.cu"] + [obj]
    subprocess.check_call([g.NVCC, *g.ARCH, "-shared", "-o", os.path.join(ROOT, "profiles", "_ab", "libdedf_trace.so"), *objs])
    print("built")
    sys.exit(0)

import torch

from diffusion_edf_b200 import FeaturedPoints, MultiscaleScoreModel, _lib as L
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
names = ["entry", "TP weights in registers, V copies issued", "pdl_wait", "row pointers, first chunk issued", "softmax max (logits from L2)",
         "softmax sum, log Z", "harmonics staged", "chunk arrived", "alpha, bias sums", "edge loop", "folds + reduced TP output",
         "V weights arrived", "linear layer + store"]
lib = L.load()
buf = (C.c_longlong * 24)()
with torch.no_grad():
    keys = model.get_key_pcd_multiscale(key)
    q = model.get_query_pcd(grasp)
    src = model.score_head.key_tensor_field.encode_sources(keys)
    n_steps = 64
    rows = [[0.5, 1e-3, 1e-3, 1.0]] * n_steps
    rows_all = model.score_head.time_rows_for(torch.full((n_steps,), 0.5, device=dev))
    dg = DenoiseGraph(model, n_poses, n_steps, src, q, False, dev)
    dg.run(T_seed.double().to(dev), src, q, rows, rows_all, None, 0)
    torch.cuda.synchronize()
    for it in range(3):
        dg.counter.zero_()
        for _ in range(20):
            dg.graph.replay()
        torch.cuda.synchronize()
        lib.dedf_vr_trace(buf)
        t = list(buf)[:13]
        
        print(f"replay batch {it}: total {t[12] - t[0]} cycles")
    for i in range(1, 13):
        print(f"   {names[i]:34s} +{t[i] - t[i - 1]:7d}   @{t[i] - t[0]:7d}")
