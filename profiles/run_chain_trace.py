"""Phase timeline of dedf_node_chain inside one CTA (clock64 stamps of thread 0, CTA 0):
    python profiles/run_chain_trace.py build      # here: profiles/_ab/libdedf_trace.so = the library with -DDEDF_CHAIN_TRACE
    DEDF_LIB=profiles/_ab/libdedf_trace.so python profiles/run_chain_trace.py [G=32] [n=16]      # on the GPU box
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
    obj = os.path.join(ROOT, "profiles", "_ab", "node_chain_trace.o")
    subprocess.check_call([g.NVCC, "-O3", "-std=c++17", "-lineinfo", *g.ARCH, "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
                           "-DDEDF_CHAIN_TRACE", "-c", os.path.join(csrc, "node_chain.cu"), "-o", obj])
    objs = [os.path.join(csrc, s.replace(".cu", ".o")) for s in g.SOURCES if s != "node_chain.cu"] + [obj]
    subprocess.check_call([g.NVCC, *g.ARCH, "-shared", "-o", os.path.join(ROOT, "profiles", "_ab", "libdedf_trace.so"), *objs])
    print("built")
    sys.exit(0)

import torch

from diffusion_edf_b200 import _lib as L, layers
from diffusion_edf_b200.block import node_tail

G = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
IRR = {16: "32x0e+16x1e+8x2e", 32: "64x0e+32x1e+16x2e"}[G]
MID = {16: "96x0e+48x1e+24x2e", 32: "192x0e+96x1e+48x2e"}[G]
dev = torch.device("cuda:0")
torch.manual_seed(0)
ln = layers.EquivariantLayerNormV2(IRR).to(dev)
ffn = layers.FeedForwardNetwork(IRR, IRR, MID).to(dev)
proj = layers.LinearRS(IRR, IRR).to(dev)
x = torch.randn(n, ln.irreps.dim if hasattr(ln, "irreps") else proj.irreps_in.dim, device=dev)
res = torch.randn_like(x)
names = ["entry", "prologue issued", "pdl_wait", "x arrived", "A tiles", "Wp/res arrived", "proj gemm", "LN stats", "LN tiles",
         "W1 arrived", "fctp_1 gemm", "gate", "W2 arrived", "fctp_2 gemm", "store drained"]
lib = L.load()
buf = (C.c_longlong * 32)()
with torch.no_grad():
    for it in range(5):
        node_tail(proj, ln, ffn, x, res)
        torch.cuda.synchronize()
        lib.dedf_chain_trace(buf)
        t = list(buf)[:15]
        print(f"run {it}: total {t[14] - t[0]} cycles")
        if it >= 3:
            for i in range(1, 15):
                print(f"   {names[i]:18s} +{t[i] - t[i - 1]:7d}   @{t[i] - t[0]:7d}")
