"""Collision-aware pre-place trajectory optimisation (diffusion_edf_b200/collision.py, csrc/collision.cu) against the
numbers the reference's own function sources produce (tests/golden/collision_golden.npz) and against the CPU oracle on larger seeded
cases.  Tolerance 1e-4 relative (fp32 sums in a different order; the neighbour SETS are discrete and must agree)."""
import os

import numpy as np
import pytest
import torch

from tests.util import assert_close

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.mark.parametrize("name", ["knn", "radius", "knn_sparse"])
def test_collision_matches_reference_code_golden(cuda, name):
    from diffusion_edf_b200 import collision as C
    from oracle import collision as OC
    from tests.golden.collision_cases import CASES, inputs
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "collision_golden.npz"))
    g = lambda k: torch.from_numpy(G[f"{name}/{k}"])                      # noqa: E731
    cfg = CASES[name]
    x, y, Ts = inputs(name)
    Ty = OC.transform_points_batched(y.expand(len(Ts), -1, 3), Ts)          # the reference transforms on the host side of _pcd_energy
    e, gr = C._pcd_energy(x.to(cuda), Ty.to(cuda), cfg["cutoff_r"], cfg["k"], cfg["eps"], True, cfg["method"])
    assert_close(e, g("energy"), TOL, "energy")
    assert_close(gr, g("grad"), TOL, "gradient (rot | trans)")
    e2, none = C._pcd_energy(x.to(cuda), Ty.to(cuda), cfg["cutoff_r"], cfg["k"], cfg["eps"], False, cfg["method"])
    assert none is None and torch.equal(e2 > 0, e > 0)
    hit = C.check_pcd_collision(x.to(cuda), Ty.to(cuda), cfg["check_r"])
    assert hit.dtype == torch.bool and np.array_equal(hit.cpu().numpy(), G[f"{name}/hit"])
    traj = C.compute_pre_place_trajectories(Ts.to(cuda), x.to(cuda), y.to(cuda), n_steps=cfg["n_steps"], dt=cfg["dt"], cutoff_r=cfg["cutoff_r"],
                                            max_num_neighbors=cfg["k"], eps=cfg["eps"], cluster_method=cfg["method"])
    assert isinstance(traj, list) and len(traj) == len(Ts) and traj[0].shape == (cfg["n_steps"], 7)
    tr = torch.stack(traj)
    assert_close(tr, g("traj"), TOL, "trajectories (end at the place poses)")
    assert torch.equal(tr[:, -1].cpu(), Ts)                                  # reverted order: last = the sampled place pose
    # the optimisation lowers the energy of every colliding pose
    e_end, _ = C._pcd_energy(x.to(cuda), OC.transform_points_batched(y.expand(len(Ts), -1, 3), tr[:, 0].cpu()).to(cuda), cfg["cutoff_r"], cfg["k"],
                             cfg["eps"], False, cfg["method"])
    assert bool((e_end[:-1] < e[:-1]).all()) and float(e_end[-1]) == 0.0


@pytest.mark.parametrize("method,k", [("knn", 16), ("knn", 100), ("radius", 20), ("radius", 1000)])
def test_collision_energy_vs_oracle_batched_clouds(cuda, method, k):
    """Per-pose grasp clouds (nPose, nY, 3), a denser scene (the k-nearest radix select runs for most queries at k = 16), Ts applied on
    the fly by the kernel (one optimisation step) against the oracle's step."""
    from diffusion_edf_b200 import collision as C
    from oracle import collision as OC
    gen = torch.Generator().manual_seed(7)
    x = torch.rand(6000, 3, generator=gen) * torch.tensor([4.0, 4.0, 1.0])
    y = (torch.rand(9, 120, 3, generator=gen) - 0.5) * 0.8
    q = torch.randn(9, 4, generator=gen)
    Ts = torch.cat([q / q.norm(dim=-1, keepdim=True), torch.rand(9, 3, generator=gen) * torch.tensor([4.0, 4.0, 1.5])], dim=-1)
    new_o, e_o = OC.optimize_once(x, y, Ts, 2e-6, 0.3, k, 0.01, method)
    new, e = C._optimize_pcd_collision_once(x.to(cuda), y.to(cuda), Ts.to(cuda), 2e-6, 0.3, k, 0.01, method)
    assert float(e_o.min()) > 0
    assert_close(e, e_o, TOL, "energy")
    assert_close(new, new_o, TOL, "updated poses")
    assert float((new_o - Ts).abs().max()) > 1e-4                            # the step did move the poses


def test_collision_edge_cases(cuda):
    from diffusion_edf_b200 import collision as C
    x = torch.rand(50, 3, device=cuda)
    y = torch.rand(4, 10, 3, device=cuda) + 10.0                            # far away: no neighbours within the cut-off
    e, g = C._pcd_energy(x, y, 0.2)
    assert e.shape == (4,) and g.shape == (4, 6) and float(e.abs().max()) == 0.0 and float(g.abs().max()) == 0.0
    assert not bool(C.check_pcd_collision(x, y, 0.2).any())
    e, g = C._pcd_energy(x[:0], y, 0.2)                                      # empty scene
    assert float(e.abs().max()) == 0.0
    e, _ = C._pcd_energy(x, x[:7].unsqueeze(0), 0.2, max_num_neighbor=100)   # fewer scene points than k, coincident points (r1 = 0)
    assert torch.isfinite(e).all() and float(e[0]) >= 7 * 0.2 / (0.001 * 0.2) * 0.999
    with pytest.raises(ValueError):
        C._pcd_energy(x, y, 0.2, cluster_method="kdtree")
    traj = C._optimize_pcd_collision_trajectory(x, y[0], torch.tensor([[1.0, 0, 0, 0, 0, 0, 0]], device=cuda).repeat(5, 1), n_steps=1, dt=0.1, cutoff_r=0.2)
    assert traj.shape == (5, 1, 7)
