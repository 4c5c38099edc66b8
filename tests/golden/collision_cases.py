"""Inputs of the collision-optimisation golden cases (tests/golden/make_golden_collision.py writes the fixture from them with the
reference's own source; tests re-create them).  Needs neither /root/reference nor a GPU."""
import torch

# scene = a table top (z = 0 slab) with a box on it; the grasped object = a small blob placed by the poses so that some poses
# penetrate the table / the box and some are clear of both.  Lengths in "rescaled" units (cut-off 0.6: several scene points per ball;
# k = 24 < the number of scene points inside the densest balls, so the k-nearest selection is exercised; dt is sized for steps of a
# few hundredths of a unit: the energies fall monotonically along the trajectories).
CASES = {
    "knn": dict(method="knn", cutoff_r=0.6, k=24, eps=0.01, dt=5e-5, n_steps=6, check_r=0.25),
    "radius": dict(method="radius", cutoff_r=0.6, k=24, eps=0.01, dt=5e-5, n_steps=5, check_r=0.25),
    "knn_sparse": dict(method="knn", cutoff_r=0.35, k=100, eps=0.01, dt=3e-4, n_steps=4, check_r=0.1),
}


def inputs(name: str):
    g = torch.Generator().manual_seed({"knn": 0, "radius": 1, "knn_sparse": 2}[name])
    n_table, n_box, n_y, n_pose = 1800, 700, 160, 7
    table = torch.rand(n_table, 3, generator=g) * torch.tensor([8.0, 8.0, 0.2]) - torch.tensor([4.0, 4.0, 0.2])
    box = torch.rand(n_box, 3, generator=g) * torch.tensor([2.0, 1.5, 1.2]) + torch.tensor([-1.0, 0.5, 0.0])
    x = torch.cat([table, box], dim=0)
    x = x[torch.randperm(len(x), generator=g)].contiguous()
    y = (torch.rand(n_y, 3, generator=g) - 0.5) * torch.tensor([0.9, 0.6, 0.7])
    q = torch.randn(n_pose, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    q = torch.where(q[:, :1] < 0, -q, q)
    p = torch.stack([torch.rand(n_pose, generator=g) * 5 - 2.5, torch.rand(n_pose, generator=g) * 5 - 2.5,
                     torch.rand(n_pose, generator=g) * 1.2 - 0.1], dim=-1)
    p[-1] = torch.tensor([0.0, 0.0, 4.0])          # one pose far above everything: zero energy, unchanged by the optimisation
    return x, y.contiguous(), torch.cat([q, p], dim=-1).contiguous()
