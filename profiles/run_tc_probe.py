"""tcgen05.mma throughput for the operand forms of the per-edge kernels (one CTA, back-to-back MMAs into one accumulator).
    python profiles/run_tc_probe.py"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusion_edf_b200 import _lib

lib = _lib.load()
lib.dedf_tc_probe.argtypes = [ctypes.c_int] * 7 + [ctypes.c_void_p, ctypes.c_void_p]
dev = torch.device("cuda:0")
out = torch.zeros(2, dtype=torch.int64, device=dev)
rows = []
for kind, kname in ((0, "tf32 K=8"),):
    for a_tmem in (0, 1):
        for N in (16, 32, 64, 128, 224):
            for n_acc, ce in ((1, 0), (4, 0), (1, 4), (1, 2), (1, 1)):
                if n_acc * N > 448:
                    continue
                reps = 480
                for _ in range(2):
                    rc = lib.dedf_tc_probe(kind, N, reps, a_tmem, 8, n_acc, ce, out.data_ptr(), None)
                    torch.cuda.synchronize()
                assert rc == 0, rc
                issue, total = out.tolist()
                rows.append({"kind": kname, "A": "tmem" if a_tmem else "smem", "N": N, "n_acc": n_acc, "commit_every": ce, "reps": reps,
                             "issue_cyc_per_mma": round(issue / reps, 1), "cyc_per_mma": round(total / reps, 1)})
print(json.dumps(rows))
for r in rows:
    print(r["kind"], "A=" + r["A"], "N=%d" % r["N"], "acc=%d" % r["n_acc"], "commit_every=%d" % r["commit_every"], r["cyc_per_mma"], file=sys.stderr)
