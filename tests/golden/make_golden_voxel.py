"""Golden vectors for voxel_filter from the REFERENCE's own function source, on the reference's own test scene (run in the
authoring container only: needs /root/reference).

`edf_interface/edf_interface/data/pcd_utils.py` cannot be imported here (plotly / torch_scatter / torch_cluster are absent),
so the `voxel_filter` function is compiled UNMODIFIED from that file's AST and executed with the real numpy / torch and a
stub `torch_scatter.scatter` (sum over an index == index_add_; that one third-party op is the part that stays unpinned).
Input: the first 6000 points (metres) and colours of edf_interface/test_data/data/demo_0/step_0/scene_pcd, 1 cm voxels,
both coordinate reductions.  Output: tests/golden/voxel_golden.npz (committed; /root/reference is never read at test time).

    python tests/golden/make_golden_voxel.py
"""
import ast
import os
import types
from typing import Tuple

import numpy as np
import torch

REF = "/root/reference/edf_interface/edf_interface/data/pcd_utils.py"
DATA = "/root/reference/edf_interface/test_data/data/demo_0/step_0/scene_pcd"


def _scatter(src, index, dim=-1, dim_size=None, reduce="sum"):
    assert reduce == "sum"
    if src.dim() == 1:
        return torch.zeros(int(dim_size), dtype=src.dtype).index_add_(0, index, src)
    return torch.zeros(int(dim_size), src.shape[1], dtype=src.dtype).index_add_(0, index.reshape(-1), src)


def load_reference_voxel_filter():
    tree = ast.parse(open(REF).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "voxel_filter")
    fn.decorator_list = []                                   # @beartype only type-checks the arguments
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"torch": torch, "np": np, "Tuple": Tuple, "torch_scatter": types.SimpleNamespace(scatter=_scatter)}
    exec(compile(mod, REF, "exec"), ns)
    return ns["voxel_filter"]


def main():
    vf = load_reference_voxel_filter()
    pts = torch.load(os.path.join(DATA, "points.pt"))[:6000].float().contiguous()
    col = torch.load(os.path.join(DATA, "colors.pt"))[:6000].float().contiguous()
    out = {"points": pts.numpy(), "colors": col.numpy(), "voxel_size": np.float32(0.01)}
    for red in ("average", "center"):
        c, f = vf(points=pts, features=col, voxel_size=0.01, coord_reduction=red)
        out[f"coord_{red}"], out[f"feat_{red}"] = c.numpy(), f.numpy()
        print(red, tuple(c.shape), tuple(f.shape))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "voxel_golden.npz"), **out)


if __name__ == "__main__":
    main()
