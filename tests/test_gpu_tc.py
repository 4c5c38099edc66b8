"""Tensor-core path (tcgen05 kind::tf32, accumulator in TMEM) against fp64 on the same seeded inputs.
n_split = 1 is plain TF32 (10-bit mantissa, ~1e-3); n_split = 3 is the 3xTF32 split the MLP kernel uses and must meet the
1e-4 bar of BASELINE.json's north_star with a wide margin (expected ~1e-6)."""
import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(16, 8), (16, 32), (32, 64), (64, 128), (128, 64), (128, 96), (240, 64), (256, 32)])
def test_tc_selftest_gemm(cuda, N, K):
    from diffusion_edf_b200 import ops
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g)
    B = torch.randn(N, K, generator=g)
    ref = A.double() @ B.double().t()
    d3 = ops.tc_selftest(A.to(cuda), B.to(cuda), 3)
    e3 = rel_err(d3, ref)
    assert e3 < 5e-6, f"3xTF32 N={N} K={K}: {e3:.3e}"
    d1 = ops.tc_selftest(A.to(cuda), B.to(cuda), 1)
    e1 = rel_err(d1, ref)
    assert 1e-6 < e1 < 5e-3, f"plain TF32 N={N} K={K}: {e1:.3e}"      # it really ran at tf32 precision


# --------------------------------------------------------------------------- the tensor-core edge MLP (dedf_edge_mlp_tc)
@pytest.mark.parametrize("fc,numel,r,E", [([32, 16, 16], 240, 3.0, 777), ([64, 32, 32], 480, 15.0, 1000), ([64, 32, 32], 240, 15.0, 129),
                                          ([32, 16, 16], 240, 3.0, 1), ([64, 32, 32], 480, 15.0, 40_000)])
def test_edge_mlp_tc_rbf(cuda, fc, numel, r, E):
    """UNet mode: GaussianRadialBasisLayerFiniteCutoff -> RadialProfile, against the oracle (radial_func.py:231-278,
    equiformer/radial_func.py:56-59) and against the fp32 CUDA-core kernel; ragged last tile, one-edge and multi-tile cases."""
    from diffusion_edf_b200 import _lib as L, layers, ops
    from oracle import encoders as enc
    from oracle import nn as ON
    torch.manual_seed(1)
    length = torch.rand(E) * r
    length[:1] = 0.0
    o_rbf = enc.GaussianRadialBasisLayerFiniteCutoff(fc[0], 0.99 * r)
    o_rad = ON.RadialProfile(fc + [numel])
    with torch.no_grad():
        ref = o_rad(o_rbf(length))
    p_rbf = layers.GaussianRadialBasisLayerFiniteCutoff(fc[0], 0.99 * r)
    p_rbf.load_state_dict(o_rbf.state_dict())
    p_rad = layers.RadialProfile(fc + [numel])
    p_rad.load_state_dict(o_rad.state_dict())
    p_rbf, p_rad = p_rbf.to(cuda), p_rad.to(cuda)
    n_dev = torch.tensor([E], dtype=torch.int32, device=cuda)
    ld = length.to(cuda)
    m, s, w = (p_rbf.mean.detach().reshape(-1), p_rbf.std_logit.detach().reshape(-1), p_rbf.weight_logit.detach().reshape(-1))
    outs = []
    for tc in (True, False):
        out = torch.full((E, numel), float("nan"), device=cuda)
        d = L.MlpDesc()
        d.mode = L.MLP_IN_RBF
        d.n_edges_dev = L.ptr(n_dev, torch.int32)
        d.length = L.ptr(ld)
        d.rbf_mean, d.rbf_std_logit, d.rbf_weight_logit = L.ptr(m), L.ptr(s), L.ptr(w)
        d.rbf_cutoff, d.rbf_offset = p_rbf.cutoff, p_rbf.offset
        p_rad.fill_desc(d, 0)
        d.out = L.ptr(out)
        assert d.W_tc[0]
        (ops.edge_mlp_tc if tc else ops.edge_mlp)(d, E)
        outs.append(out)
    e_tc, e_f32 = rel_err(outs[0], ref), rel_err(outs[1], ref)
    assert e_tc <= 1e-4, f"tensor-core MLP vs oracle: {e_tc:.3e}"
    assert e_f32 <= 1e-4
    assert rel_err(outs[0], outs[1]) <= 2e-5, "tensor-core (3xTF32) vs CUDA-core fp32 kernel"


def test_edge_mlp_tc_capacity_exceeds_edges(cuda):
    """max_edges is only a launch bound: the true edge count lives on the device and rows beyond it stay untouched."""
    from diffusion_edf_b200 import _lib as L, layers, ops
    torch.manual_seed(3)
    E, cap = 300, 5000
    p_rbf = layers.GaussianRadialBasisLayerFiniteCutoff(32, 2.97).to(cuda)
    p_rad = layers.RadialProfile([32, 16, 16, 240]).to(cuda)
    n_dev = torch.tensor([E], dtype=torch.int32, device=cuda)
    ld = (torch.rand(cap) * 3).to(cuda)
    m, s, w = (p_rbf.mean.detach().reshape(-1), p_rbf.std_logit.detach().reshape(-1), p_rbf.weight_logit.detach().reshape(-1))
    out = torch.full((cap, 240), 7.0, device=cuda)
    d = L.MlpDesc()
    d.mode = L.MLP_IN_RBF
    d.n_edges_dev = L.ptr(n_dev, torch.int32)
    d.length = L.ptr(ld)
    d.rbf_mean, d.rbf_std_logit, d.rbf_weight_logit = L.ptr(m), L.ptr(s), L.ptr(w)
    d.rbf_cutoff, d.rbf_offset = p_rbf.cutoff, p_rbf.offset
    p_rad.fill_desc(d, 0)
    d.out = L.ptr(out)
    ops.edge_mlp_tc(d, cap)
    torch.cuda.synchronize()
    assert torch.isfinite(out[:E]).all() and (out[:E] != 7.0).any()
    assert (out[E:] == 7.0).all()


@pytest.mark.parametrize("shared_time", [True, False])
def test_tensor_field_tc_vs_fp32(cuda, shared_time):
    """FIELD mode (length encoder -> per-scale pre-linear + time rows -> RadialProfile in one tensor-core launch) through the
    public MultiscaleScoreModel head: identical inputs with ops.USE_TC_MLP on / off must agree to fp32 round-off, and both
    match the oracle to 1e-4 (tests/test_gpu_model.py covers the oracle side)."""
    from diffusion_edf_b200 import FeaturedPoints, MultiscaleScoreModel, ops
    from diffusion_edf_b200.synthetic import make_poses, make_scene, model_kwargs
    torch.manual_seed(0)
    model = MultiscaleScoreModel(**model_kwargs(), deterministic=True).eval().to(cuda)
    model.use_cuda_graph = False
    x, rgb = make_scene(1500, seed=3, half_extent=12.0)
    Ts, t = make_poses(37, x, seed=3, spread=6.0)
    key = FeaturedPoints(x.to(cuda), rgb.to(cuda), torch.zeros(len(x), dtype=torch.long, device=cuda))
    grasp = FeaturedPoints(torch.zeros(8, 3, device=cuda), torch.zeros(8, 3, device=cuda), torch.zeros(8, dtype=torch.long, device=cuda))
    res = []
    with torch.no_grad():
        keys = model.get_key_pcd_multiscale(key)
        q = model.get_query_pcd(grasp)
        for tc in (True, False):
            ops.USE_TC_MLP = tc
            try:
                time = t[:1].to(cuda) if shared_time else t.to(cuda)
                res.append(model.score_head(Ts=Ts.to(cuda), key_pcd_multiscale=keys, query_pcd=q, time=time, shared_time=shared_time))
            finally:
                ops.USE_TC_MLP = True
    for a, b in zip(res[0], res[1]):
        assert rel_err(a, b) <= 2e-5
