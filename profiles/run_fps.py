"""FPS timing (CUDA events, after warm-up):  python profiles/run_fps.py [n=10000] [ratio=0.2]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusion_edf_b200 import ops
from diffusion_edf_b200.synthetic import make_scene

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
ratio = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
x, _ = make_scene(n, seed=0)
x = x.cuda()
for _ in range(3):
    idx = ops.fps(x, None, ratio)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    idx = ops.fps(x, None, ratio)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"fps n={n} m={len(idx)}: {ms * 1e3:.1f} us  ({ms * 1e3 / len(idx):.3f} us / iteration)  checksum {int(idx.sum())}")
