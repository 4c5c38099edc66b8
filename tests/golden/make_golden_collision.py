"""Golden vectors of the collision-aware trajectory optimisation from the REFERENCE'S OWN function sources (run in the authoring
container only: needs /root/reference).

`edf_interface` cannot be imported here (plotly / torch_cluster / torch_scatter / beartype-decorated data classes), so the functions
are compiled UNMODIFIED from their files' ASTs (decorators stripped: @torch.jit.script / @beartype only script / type-check) and
executed with the real torch and
  * edf_interface/data/transforms.py imported as it is (pure torch: pytorch3d's se3_exp_map, matrix_to_quaternion, quaternion ops),
  * se3._exp_map / se3._multiply and pcd_utils.transform_points extracted the same way,
  * stand-ins for the three third-party ops: torch_cluster.knn (exact top-k of the squared distances), torch_cluster.radius
    (oracle/graph.py) and torch_scatter.scatter_sum (index_add_) -- the part that stays unpinned.
Output: tests/golden/collision_golden.npz (committed; /root/reference is never read at test time).

    python tests/golden/make_golden_collision.py
"""
import ast
import importlib.util
import os
import sys
import types
from typing import List, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import collision as OC          # noqa: E402  (only its knn stand-in)
from oracle import graph as OG              # noqa: E402
from tests.golden.collision_cases import CASES, inputs      # noqa: E402

REF = "/root/reference/edf_interface/edf_interface"


def _load_module(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _extract(path, names, ns):
    tree = ast.parse(open(path).read())
    for fn in tree.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in names:
            fn.decorator_list = []
            exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def load_reference():
    transforms = _load_module(f"{REF}/data/transforms.py", "ref_transforms")
    base = {"torch": torch, "F": F, "Tuple": Tuple, "Optional": Optional, "List": List, "Union": Union}
    se3_ns = dict(base)
    se3_ns.update({k: getattr(transforms, k) for k in ("quaternion_apply", "quaternion_multiply", "normalize_quaternion", "se3_exp_map",
                                                       "matrix_to_quaternion")})
    _extract(f"{REF}/data/se3.py", ["_multiply", "_exp_map"], se3_ns)
    pcd_ns = dict(base)
    pcd_ns["quaternion_apply"] = transforms.quaternion_apply
    _extract(f"{REF}/data/pcd_utils.py", ["transform_points"], pcd_ns)

    def scatter_sum(src, index, dim=-1, dim_size=None):
        return torch.zeros(int(dim_size), dtype=src.dtype).index_add_(0, index, src)

    def radius(x, y, r, max_num_neighbors=32):
        return OG.radius(x, y, r, max_num_neighbors=max_num_neighbors)

    ns = dict(base)
    ns.update({"transforms": transforms,
               "se3": types.SimpleNamespace(_exp_map=se3_ns["_exp_map"], _multiply=se3_ns["_multiply"]),
               "pcd_utils": types.SimpleNamespace(transform_points=pcd_ns["transform_points"]),
               "torch_cluster": types.SimpleNamespace(knn=lambda x, y, k: OC.knn(x, y, k), radius=radius),
               "torch_scatter": types.SimpleNamespace(scatter_sum=scatter_sum)})
    _extract(f"{REF}/utils/collision_utils.py", ["_check_pcd_collision", "_pcd_energy", "_se3_adjoint_lie_grad",
                                                 "_optimize_pcd_collision_once", "_optimize_pcd_collision_trajectory"], ns)
    return ns


def main():
    ref = load_reference()
    out = {}
    for name, cfg in CASES.items():
        x, y, Ts = inputs(name)
        Ty = ref["pcd_utils"].transform_points(y.expand(len(Ts), -1, 3), Ts, batched_pcd=True)
        energy, grad = ref["_pcd_energy"](x=x, y=Ty, cutoff_r=cfg["cutoff_r"], max_num_neighbor=cfg["k"], eps=cfg["eps"],
                                          compute_grad=True, cluster_method=cfg["method"])
        traj = ref["_optimize_pcd_collision_trajectory"](x=x, y=y, Ts=Ts, n_steps=cfg["n_steps"], dt=cfg["dt"], cutoff_r=cfg["cutoff_r"],
                                                         max_num_neighbors=cfg["k"], eps=cfg["eps"], cluster_method=cfg["method"],
                                                         revert_order=True)
        hit = ref["_check_pcd_collision"](x=x, y=Ty, r=cfg["check_r"])
        out[f"{name}/energy"], out[f"{name}/grad"] = energy.numpy(), grad.numpy()
        out[f"{name}/traj"], out[f"{name}/hit"] = traj.detach().numpy(), hit.numpy()
        print(name, "energy", energy.numpy().round(2), "hit", hit.numpy().astype(int), "traj", tuple(traj.shape),
              "moved", float((traj[:, 0, 4:] - traj[:, -1, 4:]).norm(dim=-1).max()))
    np.savez_compressed(os.path.join(HERE, "collision_golden.npz"), **out)
    print("wrote collision_golden.npz", os.path.getsize(os.path.join(HERE, "collision_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
