"""Brute-force multi-scale radius search of the score head (query points against the 4 key scales):
    python profiles/run_radius.py [n_dst=256]          (DEDF_NO_SMEM_RADIUS=1 for the global-memory variant)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusion_edf_b200 import ops

n_dst = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
sizes = [2000, 400, 80, 16]
x_src = torch.cat([(torch.rand(n, 3) - 0.5) * 40.0 for n in sizes]).to(dev)
src_off = [0, 2000, 2400, 2480, 2496]
x_dst = ((torch.rand(n_dst, 3) - 0.5) * 40.0).to(dev)
for use_b in (True, False):
    b_src = torch.zeros(2496, dtype=torch.long, device=dev) if use_b else None
    b_dst = torch.zeros(n_dst, dtype=torch.long, device=dev) if use_b else None
    for cap in (None, n_dst * 2496):
        ov = torch.zeros(1, dtype=torch.int32, device=dev)
        fn = lambda: ops.radius_csr(x_src, x_dst, [5.0, 10.0, 20.0, None], src_off=src_off, b_src=b_src, b_dst=b_dst, capacity=cap, overflow=ov)
        for _ in range(3):
            g = fn()
        torch.cuda.synchronize()
        ops.PROFILE = {}
        for _ in range(10):
            g = fn()
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        print(f"n_dst={n_dst} batch={use_b} capacity={cap}: edges={int(g.n_edges_dev)}",
              {k: round(sum(a.elapsed_time(b) for a, b in v) / len(v) * 1e3, 1) for k, v in prof.items()}, "us")
