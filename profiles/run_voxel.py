"""Voxel filter (pre-processing) timing: CUDA path vs the CPU oracle on a 60k-point synthetic scene
(the size of the reference's raw demo scenes), 1 cm voxels.   python profiles/run_voxel.py [n=60000]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusion_edf_b200 import preprocess
from oracle.graph import voxel_filter as voxel_filter_oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000
g = torch.Generator().manual_seed(0)
p = (torch.rand(n, 3, generator=g) - 0.5) * torch.tensor([0.6, 0.6, 0.02]) + torch.tensor([0.0, 0.0, 0.3 * 0.0])
c = torch.rand(n, 3, generator=g)
pd, cd = p.cuda(), c.cuda()
for _ in range(3):
    out = preprocess.voxel_filter(pd, cd, 0.01)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    out = preprocess.voxel_filter(pd, cd, 0.01)
torch.cuda.synchronize()
gpu_ms = (time.perf_counter() - t0) / 20 * 1e3
t0 = time.perf_counter()
for _ in range(5):
    ref = voxel_filter_oracle(p, c, 0.01)
cpu_ms = (time.perf_counter() - t0) / 5 * 1e3
print(json.dumps({"workload": f"voxel_filter, {n} points, 1 cm voxels -> {out[0].shape[0]} voxels", "cuda_ms_wall_incl_2_host_reads": gpu_ms,
                  "cpu_oracle_ms": cpu_ms, "bit_exact": bool(torch.equal(out[0].cpu(), ref[0]) and torch.equal(out[1].cpu(), ref[1]))}))
