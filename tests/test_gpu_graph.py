"""Index kernels (fps / radius): BIT-EXACT against the oracle (oracle/graph.py restates torch_cluster)."""
import math

import pytest
import torch

from oracle import graph as OG

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,ratio", [(1, 1.0), (7, 0.5), (100, 0.2), (1500, 0.2), (4096, 0.1), (10_000, 0.2), (16_384, 0.05), (20_000, 0.05)])
def test_fps_bitexact(cuda, n, ratio):
    from diffusion_edf_b200 import ops
    g = torch.Generator().manual_seed(n)
    x = torch.rand(n, 3, generator=g) * 30
    ref = OG.fps(x, None, ratio)
    got = ops.fps(x.to(cuda), None, ratio).cpu()
    assert got.dtype == torch.long and got.shape == ref.shape
    assert torch.equal(got, ref), f"first mismatch at {(got != ref).nonzero()[:3].flatten().tolist()}"


@pytest.mark.parametrize("kind", ["surface", "duplicates", "flat", "line", "clustered", "all_selected", "scene"])
def test_fps_hard_clouds_bitexact(cuda, kind):
    """FPS on clouds that stress the arg-max: surface-like and clustered clouds, exact duplicates (ties -> lowest index), degenerate
    extents (a plane, a line), every point selected (ratio 1), the bench scene.  (Written for the round-2 bucketed / chunked FPS
    experiments -- bit-exact but slower than the cluster kernel, DESIGN section 8 -- and kept for the kernels that ship.)"""
    from diffusion_edf_b200 import ops
    from diffusion_edf_b200.synthetic import make_scene
    g = torch.Generator().manual_seed(17)
    n, ratio = 6000, 0.2
    if kind == "surface":
        u = torch.rand(n, 2, generator=g) * 20
        x = torch.stack([u[:, 0], u[:, 1], torch.sin(u[:, 0]) * torch.cos(u[:, 1])], dim=1)
    elif kind == "duplicates":
        x = torch.rand(n, 3, generator=g) * 10
        x[1000:2000] = x[0:1000]
        x[5000:5500] = x[2500]
    elif kind == "flat":
        x = torch.rand(n, 3, generator=g) * 10
        x[:, 2] = 3.0
    elif kind == "line":
        x = torch.zeros(n, 3)
        x[:, 0] = torch.rand(n, generator=g) * 100
    elif kind == "clustered":
        c = torch.rand(12, 3, generator=g) * 50
        x = c[torch.randint(0, 12, (n,), generator=g)] + torch.randn(n, 3, generator=g) * 0.3
    elif kind == "all_selected":
        n, ratio = 2500, 1.0
        x = torch.rand(n, 3, generator=g) * 5
    else:
        x, _ = make_scene(10_000, seed=0)
    ref = OG.fps(x, None, ratio)
    got = ops.fps(x.to(cuda), None, ratio).cpu()
    assert got.shape == ref.shape
    assert torch.equal(got, ref), f"{kind}: first mismatch at {(got != ref).nonzero()[:3].flatten().tolist()}"


def test_fps_batched_and_duplicates(cuda):
    from diffusion_edf_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.rand(300, 3, generator=g)
    x[50:60] = x[40:50]                       # exact duplicates -> ties
    b = torch.cat([torch.zeros(120), torch.ones(180)]).long()
    ref = OG.fps(x, b, 0.25)
    got = ops.fps(x.to(cuda), b.to(cuda), 0.25).cpu()
    assert torch.equal(got, ref)


@pytest.mark.parametrize("n_src,n_dst,r,max_nb", [(0 + 50, 0 + 1, 0.5, 32), (2000, 256, 5.0, 1000), (10_000, 2000, 3.0, 1000),
                                                 (500, 500, 8.0, 16)])
def test_radius_bitexact(cuda, n_src, n_dst, r, max_nb):
    from diffusion_edf_b200 import ops
    g = torch.Generator().manual_seed(n_src + n_dst)
    x = torch.rand(n_src, 3, generator=g) * 30
    y = torch.rand(n_dst, 3, generator=g) * 30
    ref = OG.radius(x, y, r, None, None, max_nb)
    got = ops.radius(x.to(cuda), y.to(cuda), r, max_num_neighbors=max_nb).cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert torch.equal(got, ref)


def test_radius_batch_and_empty(cuda):
    from diffusion_edf_b200 import ops
    g = torch.Generator().manual_seed(9)
    x = torch.rand(400, 3, generator=g) * 10
    y = torch.rand(60, 3, generator=g) * 10
    bx = torch.cat([torch.zeros(150), torch.ones(250)]).long()
    by = torch.cat([torch.zeros(20), torch.ones(40)]).long()
    ref = OG.radius(x, y, 2.0, bx, by, 64)
    got = ops.radius(x.to(cuda), y.to(cuda), 2.0, bx.to(cuda), by.to(cuda), 64).cpu()
    assert torch.equal(got, ref)
    # nothing within reach -> empty edge list, still a valid CSR
    far = ops.radius_csr(x.to(cuda), (y + 1000).to(cuda), [2.0])
    assert far.n_edges == 0 and int(far.row_ptr[-1]) == 0


def test_radius_graph_bitexact(cuda):
    from diffusion_edf_b200 import ops
    g = torch.Generator().manual_seed(11)
    x = torch.rand(2000, 3, generator=g) * 25
    ref = OG.radius_graph(x, 3.0, None, loop=False, max_num_neighbors=1000)
    got = ops.radius_graph(x.to(cuda), 3.0, max_num_neighbors=1000).cpu()
    assert torch.equal(got, ref)
    # truncation below the degree: first max_nb in index order survive
    ref = OG.radius_graph(x, 6.0, None, loop=False, max_num_neighbors=8)
    got = ops.radius_graph(x.to(cuda), 6.0, max_num_neighbors=8).cpu()
    assert torch.equal(got, ref)


def test_multiscale_csr_matches_per_scale(cuda):
    """One launch over 4 source clouds == the reference's per-scale radius + all-pairs concatenation
    (multiscale_tensor_field.py:208-247), up to the order inside the all-pairs scale."""
    from diffusion_edf_b200 import ops
    g = torch.Generator().manual_seed(3)
    clouds = [torch.rand(n, 3, generator=g) * 40 for n in (2000, 400, 80, 16)]
    y = torch.rand(64, 3, generator=g) * 40
    radii = [5.0, 10.0, 20.0, None]
    off = [0]
    for c in clouds:
        off.append(off[-1] + len(c))
    csr = ops.radius_csr(torch.cat(clouds).to(cuda), y.to(cuda), radii, src_off=off)
    rp = csr.row_ptr.cpu().long()
    es, ed = csr.edge_src.cpu().long(), csr.edge_dst.cpu().long()
    for s, (c, r) in enumerate(zip(clouds, radii)):
        lo, hi = int(rp[s * 64]), int(rp[(s + 1) * 64])
        if r is None:
            ref_dst, ref_src = torch.meshgrid(torch.arange(64), torch.arange(len(c)), indexing="ij")
            ref = torch.stack([ref_dst.reshape(-1), ref_src.reshape(-1)])
        else:
            ref = OG.radius(c, y, r, None, None, 1000)
        assert torch.equal(ed[lo:hi], ref[0]) and torch.equal(es[lo:hi] - off[s], ref[1]), f"scale {s}"


def test_pool_exclusions(cuda):
    """FpsPool drops (src == idx[dst]) after the radius search (connectivity.py:64-70); the un-pool graph is the
    same edge set with the roles swapped."""
    from diffusion_edf_b200 import ops
    g = torch.Generator().manual_seed(21)
    x = torch.rand(1200, 3, generator=g) * 20
    idx = OG.fps(x, None, 0.2)
    xd = x[idx]
    e = OG.radius(x, xd, 3.0, None, None, 1000)
    keep = idx[e[0]] != e[1]
    ref_dst, ref_src = e[0][keep], e[1][keep]
    csr = ops.radius_csr(x.to(cuda), xd.to(cuda), [3.0], excl_mode=1, excl=idx.to(cuda))
    assert torch.equal(csr.edge_dst.cpu().long(), ref_dst) and torch.equal(csr.edge_src.cpu().long(), ref_src)
    rev = ops.radius_csr(xd.to(cuda), x.to(cuda), [3.0], excl_mode=3, excl=idx.to(cuda))
    a = set(zip(rev.edge_dst.cpu().tolist(), rev.edge_src.cpu().tolist()))
    bset = set(zip(ref_src.tolist(), ref_dst.tolist()))
    assert a == bset


@pytest.mark.parametrize("n_src,n_dst,r,max_nb,excl_mode", [(5000, 700, 2.5, 1000, 0), (5000, 5000, 2.0, 1000, 2), (3000, 200, 30.0, 1000, 0),
                                                           (4000, 300, 6.0, 24, 0), (2500, 500, 3.0, 1000, 1)])
def test_grid_hash_equals_brute_force(cuda, n_src, n_dst, r, max_nb, excl_mode):
    """The grid-hash kernels return element-for-element the CSR of the ordered brute-force kernel (negative coordinates,
    truncation, exclusions, batches, and the > 1024-hit fallback at r = 30)."""
    from diffusion_edf_b200 import ops
    g = torch.Generator().manual_seed(n_src + n_dst)
    x = (torch.rand(n_src, 3, generator=g) - 0.5) * 40
    y = x[:n_dst].clone() if excl_mode == 2 else (torch.rand(n_dst, 3, generator=g) - 0.5) * 40
    if excl_mode == 2:
        x = y
    bx = (torch.arange(len(x)) >= len(x) // 3).long()
    by = (torch.arange(len(y)) >= len(y) // 3).long()
    excl = torch.randint(0, len(x), (len(y),), generator=g) if excl_mode == 1 else None
    kw = dict(b_src=bx.to(cuda), b_dst=by.to(cuda), excl_mode=excl_mode, excl=None if excl is None else excl.to(cuda), max_num_neighbors=max_nb)
    old = ops.GRID_MIN_SOURCES
    try:
        ops.GRID_MIN_SOURCES = 1 << 30
        brute = ops.radius_csr(x.to(cuda), y.to(cuda), [r], **kw)
        ops.GRID_MIN_SOURCES = 1
        grid = ops.radius_csr(x.to(cuda), y.to(cuda), [r], **kw)
    finally:
        ops.GRID_MIN_SOURCES = old
    assert grid.n_edges == brute.n_edges and brute.n_edges > 0
    assert torch.equal(grid.row_ptr, brute.row_ptr)
    assert torch.equal(grid.edge_src, brute.edge_src) and torch.equal(grid.edge_dst, brute.edge_dst)
