"""edge_tp_lin micro-benchmark + phase timeline of CTA 0:  python profiles/run_tp_lin.py G n_src n_dst deg"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusion_edf_b200 import _lib as L
from diffusion_edf_b200 import ops
from diffusion_edf_b200.irreps import Irreps
from diffusion_edf_b200.layers import GraphAttention

G = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n_src = int(sys.argv[2]) if len(sys.argv) > 2 else 2496
n_dst = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
deg = int(sys.argv[4]) if len(sys.argv) > 4 else 42
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
logits = torch.empty(E, 4, device=dev); v = torch.empty(E, F, device=dev); val = torch.empty(E, F, device=dev)
lib = L.load()
lib.dedf_tp_lin_set_debug.argtypes = [ctypes.c_void_p]


def act():
    ops.edge_tp_lin(G, L.EPI_ACT, msg, None, False, g, sh, w, ga.sep_act.numel, p["W0"], p["W1"], p["W2"], p["b0"],
                    alpha_dot=p["alpha_dot"], edge_logit=None, logits=logits, out=v)


def lin():
    ops.edge_tp_lin(G, L.EPI_LIN, v, None, True, g, sh, p["wv"], 0, p["V0"], p["V1"], p["V2"], p["vb"], out=val)


def act_tc():
    ops.edge_tp_act_tc(G, msg, None, g, sh, w, ga.sep_act.numel, p["Wtc"], p["b0"], p["alpha_dot"], None, logits, v)


from diffusion_edf_b200 import layers
w_p = w[:, layers.tp_act_w_perm(G).to(dev)].contiguous()
msg_d = torch.randn(n_dst, F, device=dev)


def act_tc_perm():
    ops.edge_tp_act_tc(G, msg, None, g, sh, w_p, ga.sep_act.numel, p["Wtc"], p["b0"], p["alpha_dot"], None, logits, v, w_perm=True)


def act_tc_perm_dst():
    ops.edge_tp_act_tc(G, msg, msg_d, g, sh, w_p, ga.sep_act.numel, p["Wtc"], p["b0"], p["alpha_dot"], None, logits, v, w_perm=True)


def act_dst():
    ops.edge_tp_lin(G, L.EPI_ACT, msg, msg_d, False, g, sh, w, ga.sep_act.numel, p["W0"], p["W1"], p["W2"], p["b0"],
                    alpha_dot=p["alpha_dot"], edge_logit=None, logits=logits, out=v)


for name, fn in (("ACT", act), ("ACT+dst", act_dst), ("ACT_TC", act_tc), ("ACT_TC perm", act_tc_perm), ("ACT_TC perm+dst", act_tc_perm_dst), ("LIN", lin)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    dbg = torch.zeros(64, dtype=torch.int64, device=dev)
    lib.dedf_tp_lin_set_debug(dbg.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.dedf_tp_lin_set_debug(None)
    d = [z for z in dbg.cpu().tolist() if z]
    if name.startswith("ACT_TC"):
        lib.dedf_tp_act_tc_set_debug.argtypes = [ctypes.c_void_p]
        d2 = torch.zeros(8, dtype=torch.int64, device=dev)
        lib.dedf_tp_act_tc_set_debug(d2.data_ptr()); fn(); torch.cuda.synchronize(); lib.dedf_tp_act_tc_set_debug(None)
        print("    MMA issuer CTA0 [total, wait A, wait W, wait acc, issue, tiles]:", d2.cpu().tolist()[:6])
    print(f"{name} G={G} E={E}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us;  CTA0 stamps (cycles, per tile: start, staged, CG done, GEMM done, O written):")
    print("   ", [z - d[0] for z in d][:31])
