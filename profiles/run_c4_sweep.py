"""BASELINE config C4: K1 (fused gather -> depthwise CG TP -> x alpha -> segment reduce) on self-graphs built with the
grid-hash radius kernel, N in {1k,3k,10k,30k,100k}, uniform density giving average degree 32 at r = 1.
Prints one JSON line per N:   python profiles/run_c4_sweep.py"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusion_edf_b200 import ops

dev = torch.device("cuda:0")
PEAK = 6540.5
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
rho = 32 / (4 * math.pi / 3)
flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
for N in (1000, 3000, 10_000, 30_000, 100_000):
    g = torch.Generator().manual_seed(N)
    side = (N / rho) ** (1 / 3)
    x = (torch.rand(N, 3, generator=g) * side).to(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    csr = ops.radius_csr(x, x, [1.0], excl_mode=2, max_num_neighbors=1001)
    e1.record(); torch.cuda.synchronize()
    t_graph = e0.elapsed_time(e1)
    E = csr.n_edges
    length, sh9, _ = ops.edge_geom(x, x, csr)
    sh = torch.zeros(E, 12, device=dev); sh[:, :9] = sh9
    feat = torch.randn(N, 240, device=dev)
    w = torch.randn(E, 480, device=dev) / math.sqrt(32.0)
    alpha = torch.rand(E, 4, device=dev)
    times = []
    for it in range(3 + 10):
        flush.fill_(float(it))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = ops.edge_tp_reduce(32, feat, csr.row_ptr, csr.edge_src, sh, w, alpha)
        b.record(); torch.cuda.synchronize()
        if it >= 3:
            times.append(a.elapsed_time(b))
    ms = sum(times) / len(times)
    alg = E * 4 * (480 + 9 + 4 + 1) + N * 4 * 240 + N * 4 * 1568 + 4 * (N + 1)
    print(json.dumps({"workload": "C4", "N": N, "E": E, "avg_degree": E / N, "graph_build_ms": t_graph, "k1_ms": ms,
                      "algorithmic_bytes": alg, "achieved_gbs": alg / ms / 1e6, "frac_of_measured_hbm_peak": alg / ms / 1e6 / PEAK,
                      "finite": bool(torch.isfinite(out).all())}))
