"""Pre-processing: the CUDA voxel filter against the golden vectors produced by the reference's own
voxel_filter on its test scene, and against the oracle on random clouds.  Bit-exact (integer / index work plus sums taken in
the same order as a sequential scatter)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_voxel_filter_reference_golden(cuda):
    from diffusion_edf_b200 import preprocess
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "voxel_golden.npz"))
    p, c = torch.tensor(g["points"]).to(cuda), torch.tensor(g["colors"]).to(cuda)
    for red in ("average", "center"):
        co, fe = preprocess.voxel_filter(p, c, float(g["voxel_size"]), red)
        assert torch.equal(co.cpu(), torch.tensor(g[f"coord_{red}"])), red
        assert torch.equal(fe.cpu(), torch.tensor(g[f"feat_{red}"])), red


@pytest.mark.parametrize("n,vs,F", [(1, 0.1, 3), (50_000, 0.01, 3), (20_000, 0.05, 7), (3000, 10.0, 1)])
def test_voxel_filter_vs_oracle(cuda, n, vs, F):
    from diffusion_edf_b200 import preprocess
    from oracle.graph import voxel_filter
    gen = torch.Generator().manual_seed(n)
    p = (torch.rand(n, 3, generator=gen) - 0.5) * torch.tensor([0.6, 0.6, 0.3])
    f = torch.rand(n, F, generator=gen)
    for red in ("average", "center"):
        co_o, fe_o = voxel_filter(p, f, vs, red)
        co, fe = preprocess.voxel_filter(p.to(cuda), f.to(cuda), vs, red)
        assert co.shape == co_o.shape
        assert torch.equal(co.cpu(), co_o) and torch.equal(fe.cpu(), fe_o)
    with pytest.raises(ValueError):
        preprocess.voxel_filter(p.to(cuda), f.to(cuda), vs, "median")


def test_downsample_rescale_feeds_the_model_types(cuda):
    from diffusion_edf_b200 import FeaturedPoints, preprocess
    gen = torch.Generator().manual_seed(0)
    p = torch.rand(5000, 3, generator=gen) * 0.3
    c = torch.rand(5000, 3, generator=gen)
    fp = preprocess.downsample_and_rescale(p.to(cuda), c.to(cuda), voxel_size=0.01, rescale_factor=100.0)
    assert isinstance(fp, FeaturedPoints) and fp.x.shape[1] == 3 and fp.f.shape == fp.x.shape and fp.b.dtype == torch.long
    assert float(fp.x.max()) <= 30.0 + 1e-3 and float(fp.x.max()) > 20.0       # metres -> centimetres
