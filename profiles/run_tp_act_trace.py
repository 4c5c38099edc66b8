"""Where a producer warp of edge_tp_act_tc spends its chunk loop (clock64 segments of warp 0, CTA 0):
    python profiles/run_tp_act_trace.py build       # here: profiles/_ab/libdedf_trace.so = the library with -DDEDF_TA_TRACE
    DEDF_LIB=profiles/_ab/libdedf_trace.so python profiles/run_tp_act_trace.py [G=32] [n_dst=2048] [deg=42] [f16=1]      # on the GPU box
"""
import ctypes
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
    obj = os.path.join(ROOT, "profiles", "_ab", "tc_tplin_trace.o")
    subprocess.check_call([g.NVCC, "-O3", "-std=c++17", "-lineinfo", *g.ARCH, "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
                           "-DDEDF_TA_TRACE", "-c", os.path.join(csrc, "tc_tplin.cu"), "-o", obj])
    objs = [os.path.join(csrc, s.replace(".cu", ".o")) for s in g.SOURCES if s != "tc_tplin.cu"] + [obj]
    subprocess.check_call([g.NVCC, *g.ARCH, "-shared", "-o", os.path.join(ROOT, "profiles", "_ab", "libdedf_trace.so"), *objs])
    print("built")
    sys.exit(0)

import torch

from diffusion_edf_b200 import _lib as L
from diffusion_edf_b200 import layers, ops
from diffusion_edf_b200.irreps import Irreps
from diffusion_edf_b200.layers import GraphAttention

G = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n_dst = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
deg = int(sys.argv[3]) if len(sys.argv) > 3 else 42
f16 = (int(sys.argv[4]) if len(sys.argv) > 4 else 1) != 0
n_src = 2496
dev = torch.device("cuda:0")
torch.manual_seed(0)
irr = Irreps((2 * G, G, G // 2))
ga = GraphAttention(irr, irr, [64, 32, 32], 4).to(dev)
E = n_dst * deg
row_ptr = (torch.arange(n_dst + 1) * deg).int().to(dev)
edge_src = torch.randint(0, n_src, (E,), dtype=torch.int32).to(dev)
edge_dst = torch.arange(n_dst, dtype=torch.int32).repeat_interleave(deg).to(dev)
g = ops.Csr(row_ptr, edge_src, edge_dst, row_ptr[-1:], E, n_dst, 1)
F = irr.dim
msg = torch.randn(n_src, F, device=dev)
sh = torch.randn(E, 9, device=dev)
w = torch.randn(E, ga.sep_act.numel, device=dev) * 0.1
p = ga.packed()
logits = torch.empty(E, 4, device=dev); v = torch.empty(E, F, device=dev)
w_p = w[:, layers.tp_act_w_perm(G).to(dev)].contiguous()
lib = L.load()
lib.dedf_tp_act_tc_set_debug.argtypes = [ctypes.c_void_p]


def fn():
    ops.edge_tp_act_tc(G, msg, None, g, sh, w_p, ga.sep_act.numel, p["Wtc16"] if f16 else p["Wtc"], p["b0"], p["alpha_dot"], None, logits, v,
                       w_perm=True, f16=f16)


for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    fn()
e1.record()
torch.cuda.synchronize()
d = torch.zeros(32, dtype=torch.int64, device=dev)
lib.dedf_tp_act_tc_set_debug(d.data_ptr()); fn(); torch.cuda.synchronize(); lib.dedf_tp_act_tc_set_debug(None)
d = d.cpu().tolist()
tiles = max(1, d[5]); nch = G // 4
print(f"G={G} E={E} {'fp16' if f16 else 'tf32'} split: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us, {tiles} tiles in CTA 0")
print("MMA issuer [total, wait A, wait W, wait acc, issue]:", d[:5], " per chunk:", [round(x / tiles / nch) for x in d[:5]])
names = ["tile head (index loads, first gathers)", "x slice LDS + CG math", "next w loads + wait emptyA", "hi/lo stores", "fence + arrive",
         "stage next x slice", "producer barrier"]
for k, nm in enumerate(names):
    per = d[8 + k] / tiles / (1 if k == 0 else nch)
    print(f"   {nm:42s} {d[8 + k]:10d}   {per:8.0f} cycles per {'tile' if k == 0 else 'chunk'}")
print(f"   epilogue warp 8: wait for the accumulator {d[16] / tiles:8.0f}, epilogue {d[17] / tiles:8.0f} cycles per tile")
