"""K1 on config C4 (N nodes, degree 32) for ncu captures:  python profiles/run_k1.py [N] [reps] [variant: tma|ldg]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusion_edf_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
variant = sys.argv[3] if len(sys.argv) > 3 else "tma"
if len(sys.argv) > 4 and sys.argv[4] == "nopersist":
    ops.K1_PERSIST_BYTES = None            # A/B: without the L2 access-policy window on the gathered table
dev = torch.device("cuda:0")
deg, E = 32, N * 32
g = torch.Generator().manual_seed(0)
row_ptr = (torch.arange(N + 1) * deg).int().to(dev)
edge_src = torch.randint(0, N, (E,), generator=g, dtype=torch.int32).to(dev)
x = torch.randn(N, 240, device=dev)
sh = torch.zeros(E, 12 if variant == "tma" else 9, device=dev)
sh[:, :9] = torch.randn(E, 9, device=dev)
w = torch.randn(E, 480, device=dev) * 0.1
alpha = torch.rand(E, 4, device=dev)
flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
for i in range(reps):
    flush.fill_(float(i))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ops.edge_tp_reduce(32, x, row_ptr, edge_src, sh, w, alpha)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    alg = E * 4 * (480 + 9 + 4 + 1) + N * 4 * 240 + N * 4 * 1568 + 4 * (N + 1)
    print(f"{variant} N={N} rep {i}: {ms:.3f} ms  {alg / ms / 1e6:.1f} GB/s algorithmic")
