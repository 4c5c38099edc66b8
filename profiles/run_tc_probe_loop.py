"""The MLP kernel's MMA issue loop in isolation (operands resident in shared memory): cycles per K = 8 chunk (3 MMAs) for one issuing
thread vs a warp-uniform loop with an elected issuer, and for commits every 1 / 4 / 16 chunks.   python profiles/run_tc_probe_loop.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusion_edf_b200 import _lib

lib = _lib.load()
lib.dedf_tc_probe_loop.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda:0")
for mode, name in ((0, "one thread (if lane == 0)"), (1, "whole warp + elect.sync"), (2, "warp + elect, 3 accumulators")):
    for N in (64, 128):
        for cps in (1, 16):
            for _ in range(2):
                rc = lib.dedf_tc_probe_loop(mode, N, 16, cps, 20, out.data_ptr(), None)
                torch.cuda.synchronize()
            assert rc == 0, rc
            issue, total = out.tolist()
            print(f"{name:28s} N={N:3d} commit every {cps:2d} chunks: issue {issue / 320:7.1f} cycles/chunk, complete {total / 320:7.1f} cycles/chunk")
