"""tcgen05.mma (kind::tf32, M = 128, K = 8, both operands in shared memory) cycles per MMA for the no-swizzle chunk-major operand
layout the kernels use against K-major swizzled layouts.   python profiles/run_tc_probe_swz.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusion_edf_b200 import _lib

lib = _lib.load()
lib.dedf_tc_probe_swz.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda:0")
for N in (16, 32, 48, 64, 128, 240, 256):
    row = []
    for layout, name in ((0, "none"), (2, "sw128"), (4, "sw64"), (6, "sw32")):
        reps = 512
        for _ in range(2):
            rc = lib.dedf_tc_probe_swz(N, reps, layout, out.data_ptr(), None)
            torch.cuda.synchronize()
        assert rc == 0, rc
        issue, total = out.tolist()
        row.append(f"{name}: {total / reps:6.1f} (issue {issue / reps:5.1f})")
    print(f"N={N:3d}  cycles per MMA  " + "   ".join(row))

lib.dedf_tc_probe_issue.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p, ctypes.c_void_p]
for N in (32, 48, 128, 240):
    row = []
    for variant, name in ((1, "1 thread, desc += step"), (2, "warp-uniform, elected lane"), (3, "1 thread, desc rebuilt")):
        reps = 512
        for _ in range(2):
            rc = lib.dedf_tc_probe_issue(variant, N, reps, out.data_ptr(), None)
            torch.cuda.synchronize()
        assert rc == 0, rc
        issue, total = out.tolist()
        row.append(f"{name}: {total / reps:6.1f} (issue {issue / reps:5.1f})")
    print(f"N={N:3d}  varying descriptors, cycles per MMA  " + "   ".join(row))
