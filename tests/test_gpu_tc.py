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
    for tc, f16 in ((True, False), (False, False), (True, True)):
        out = torch.full((E, numel), float("nan"), device=cuda)
        d = L.MlpDesc()
        d.mode = L.MLP_IN_RBF
        d.n_edges_dev = L.ptr(n_dev, torch.int32)
        d.length = L.ptr(ld)
        d.rbf_mean, d.rbf_std_logit, d.rbf_weight_logit = L.ptr(m), L.ptr(s), L.ptr(w)
        d.rbf_cutoff, d.rbf_offset = p_rbf.cutoff, p_rbf.offset
        p_rad.fill_desc(d, 0, f16=f16)
        d.out = L.ptr(out)
        assert d.W_tc[0] and d.tc_f16 == int(f16)
        (ops.edge_mlp_tc if tc else ops.edge_mlp)(d, E)
        outs.append(out)
    e_tc, e_f32, e_f16 = rel_err(outs[0], ref), rel_err(outs[1], ref), rel_err(outs[2], ref)
    assert e_tc <= 1e-4, f"tensor-core MLP (tf32 split) vs oracle: {e_tc:.3e}"
    assert e_f16 <= 1e-4, f"tensor-core MLP (fp16 split) vs oracle: {e_f16:.3e}"
    assert e_f32 <= 1e-4
    assert rel_err(outs[0], outs[1]) <= 2e-5, "tensor-core (3xTF32) vs CUDA-core fp32 kernel"
    assert rel_err(outs[2], outs[1]) <= 2e-5, f"tensor-core (3xFP16) vs CUDA-core fp32 kernel: {rel_err(outs[2], outs[1]):.3e}"


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
        for tc, f16 in ((True, True), (False, True), (True, False)):
            ops.USE_TC_MLP, ops.MLP_F16 = tc, f16
            try:
                time = t[:1].to(cuda) if shared_time else t.to(cuda)
                res.append(model.score_head(Ts=Ts.to(cuda), key_pcd_multiscale=keys, query_pcd=q, time=time, shared_time=shared_time))
            finally:
                ops.USE_TC_MLP, ops.MLP_F16 = True, True
    for other in (res[0], res[2]):          # fp16 split, tf32 split   vs   the CUDA-core fp32 MLP
        for a, b in zip(other, res[1]):
            assert rel_err(a, b) <= 2e-5


# --------------------------------------------------------------------------- attention logits + gated values (dedf_edge_tp_act_tc)
@pytest.mark.parametrize("G,E,use_dst,use_logit", [(32, 1, True, True), (32, 333, True, True), (32, 64, False, False),
                                                   (16, 777, True, False), (16, 31, False, True), (32, 21_001, True, True),
                                                   (16, 30_000, True, True)])
@pytest.mark.parametrize("f16", [False, True], ids=["tf32x3", "f16x3"])
def test_edge_tp_act_tc(cuda, G, E, use_dst, use_logit, f16):
    """gather -> depthwise TP -> [sep_alpha | sep_act.lin] -> SmoothLeakyReLU.alpha_dot / Gate (graph_attention.py:231-246)
    with the linear layer on the tensor cores, against the oracle arithmetic (small cases) and the fp32 CUDA-core kernel
    (every case): one edge, ragged last tile, several tiles per CTA (both TMEM accumulator buffers and ring wrap-around); with
    the tf32 hi / lo split and with the fp16 hi / lo split (kind::f16, two chunks per operand stage)."""
    from diffusion_edf_b200 import _lib as L, layers, ops
    from oracle import model as OM
    from oracle import nn as ON
    from oracle import so3
    from oracle.irreps import Irreps as OIrreps
    IRR = {32: "64x0e+32x1e+16x2e", 16: "32x0e+16x1e+8x2e"}
    gen = torch.Generator().manual_seed(G * 7 + E)
    torch.manual_seed(G * 7 + E)
    irr = OIrreps(IRR[G])
    F, numel = irr.dim, 15 * G
    n_src, n_dst = 97, 41
    es = torch.randint(0, n_src, (E,), generator=gen)
    ed = torch.randint(0, n_dst, (E,), generator=gen).sort().values
    sh = so3.spherical_harmonics(2, torch.randn(E, 3, generator=gen))
    msg_src = torch.randn(n_src, F, generator=gen)
    msg_dst = torch.randn(n_dst, F, generator=gen) if use_dst else None
    w = torch.randn(E, numel, generator=gen) / 3.0 ** 0.5
    edge_logit = -torch.rand(E, generator=gen) * 3 if use_logit else None
    oga = OM.GraphAttentionMLP2(irr, OIrreps("1x0e+1x1e+1x2e"), irr, [32, 16, 16], 4)
    with torch.no_grad():
        for prm in oga.parameters():
            if prm.abs().sum() == 0:
                prm.uniform_(-0.5, 0.5)
    pga = layers.GraphAttention(IRR[G], IRR[G], [32, 16, 16], 4)
    pga.load_state_dict(oga.state_dict())
    pga = pga.to(cuda)
    p = pga.packed()
    row_ptr = torch.zeros(n_dst + 1, dtype=torch.long)
    row_ptr[1:] = torch.bincount(ed, minlength=n_dst).cumsum(0)
    rp = row_ptr.int().to(cuda)
    csr = ops.Csr(rp, es.int().to(cuda), ed.int().to(cuda), rp[-1:], E, n_dst, 1)
    dv = lambda t: None if t is None else t.to(cuda)
    logits = torch.full((E, 4), float("nan"), device=cuda)
    v = torch.full((E, F), float("nan"), device=cuda)
    Wtc = p["Wtc16"] if f16 else p["Wtc"]
    ops.edge_tp_act_tc(G, dv(msg_src), dv(msg_dst), csr, dv(sh), dv(w), numel, Wtc, p["b0"], p["alpha_dot"], dv(edge_logit), logits, v, f16=f16)
    logits2 = torch.empty(E, 4, device=cuda)
    v2 = torch.empty(E, F, device=cuda)
    ops.edge_tp_lin(G, L.EPI_ACT, dv(msg_src), dv(msg_dst), False, csr, dv(sh), dv(w), numel, p["W0"], p["W1"], p["W2"], p["b0"],
                    alpha_dot=p["alpha_dot"], edge_logit=dv(edge_logit), logits=logits2, out=v2)
    assert torch.isfinite(logits).all() and torch.isfinite(v).all()
    assert rel_err(logits, logits2) < 2e-5, f"logits vs fp32 kernel: {rel_err(logits, logits2):.3e}"
    assert rel_err(v, v2) < 2e-5, f"values vs fp32 kernel: {rel_err(v, v2):.3e}"
    if E <= 1000:
        with torch.no_grad():
            message = msg_src[es] + (msg_dst[ed] if use_dst else 0)
            d1 = oga.sep_act.dtp(message, sh, w)
            la = oga.sep_alpha(d1).reshape(E, 4, -1)
            v_ref = oga.sep_act.gate(oga.sep_act.lin(d1))
            la = oga.c_slrelu * ON.smooth_leaky_relu(la)
            logit_ref = torch.einsum("ehk,hk->eh", la, oga.alpha_dot.squeeze(0)) + (edge_logit[:, None] if use_logit else 0)
        assert rel_err(logits, logit_ref) < 1e-4, f"logits vs oracle: {rel_err(logits, logit_ref):.3e}"
        assert rel_err(v, v_ref) < 1e-4, f"values vs oracle: {rel_err(v, v_ref):.3e}"
    # same launch twice: bit-identical; per-edge weights handed over in the kernel's chunk-major column order: bit-identical too
    logits3 = torch.empty(E, 4, device=cuda)
    v3 = torch.empty(E, F, device=cuda)
    ops.edge_tp_act_tc(G, dv(msg_src), dv(msg_dst), csr, dv(sh), dv(w), numel, Wtc, p["b0"], p["alpha_dot"], dv(edge_logit), logits3, v3, f16=f16)
    assert torch.equal(logits, logits3) and torch.equal(v, v3)
    w_perm = w[:, layers.tp_act_w_perm(G)].contiguous()
    logits3.fill_(float("nan")); v3.fill_(float("nan"))
    ops.edge_tp_act_tc(G, dv(msg_src), dv(msg_dst), csr, dv(sh), dv(w_perm), numel, Wtc, p["b0"], p["alpha_dot"], dv(edge_logit), logits3, v3,
                       w_perm=True, f16=f16)
    assert torch.equal(logits, logits3) and torch.equal(v, v3)


def test_full_forward_tc_attention_on_off(cuda):
    """Whole MultiscaleScoreModel.forward (UNet encoder with G = 16 and G = 32 blocks, tensor field, score head) with the
    attention block's linear layer on the tensor cores + chunk-major radial weights (ops.USE_TC_TPACT) and with the fp32
    CUDA-core kernel + reference column order: same scores to fp32 round-off."""
    from diffusion_edf_b200 import FeaturedPoints, MultiscaleScoreModel, ops
    from diffusion_edf_b200.synthetic import make_poses, make_scene, model_kwargs
    torch.manual_seed(0)
    model = MultiscaleScoreModel(**model_kwargs(), deterministic=True).eval().to(cuda)
    model.use_cuda_graph = False
    x, rgb = make_scene(2500, seed=5, half_extent=12.0)
    Ts, t = make_poses(19, x, seed=5, spread=6.0)
    key = FeaturedPoints(x.to(cuda), rgb.to(cuda), torch.zeros(len(x), dtype=torch.long, device=cuda))
    grasp = FeaturedPoints(torch.zeros(8, 3, device=cuda), torch.zeros(8, 3, device=cuda), torch.zeros(8, dtype=torch.long, device=cuda))
    res = []
    with torch.no_grad():
        for tc in (True, False):
            ops.USE_TC_TPACT = tc
            try:
                n0 = ops.LAUNCHES
                (ang, lin), _ = model(Ts.to(cuda), t.to(cuda), key, grasp)
                res.append((ang.clone(), lin.clone()))
            finally:
                ops.USE_TC_TPACT = True
    for a, b in zip(res[0], res[1]):
        assert torch.isfinite(a).all()
        assert rel_err(a, b) <= 5e-5, rel_err(a, b)


def test_full_forward_operand_splits_and_early_query_stream(cuda, monkeypatch):
    """Whole forward through the CUDA-graph path (a) with the fp16 hi / lo operand split of the two tcgen05 kernels (default) against
    the tf32 split (ops.MLP_F16 / ops.TPACT_F16 off): same scores to fp32 round-off; (b) with the query model + time embedding on
    their own stream under the key encoder (default) against the plain stream order (DEDF_EARLY_QUERY=0): the SAME kernels on the
    same data, so the scores must be EQUAL."""
    from diffusion_edf_b200 import FeaturedPoints, MultiscaleScoreModel, ops
    from diffusion_edf_b200.synthetic import make_poses, make_scene, model_kwargs
    torch.manual_seed(0)
    model = MultiscaleScoreModel(**model_kwargs(), deterministic=True).eval().to(cuda)
    x, rgb = make_scene(2500, seed=6, half_extent=12.0)
    Ts, t = make_poses(23, x, seed=6, spread=6.0)
    key = FeaturedPoints(x.to(cuda), rgb.to(cuda), torch.zeros(len(x), dtype=torch.long, device=cuda))
    grasp = FeaturedPoints(torch.zeros(8, 3, device=cuda), torch.zeros(8, 3, device=cuda), torch.zeros(8, dtype=torch.long, device=cuda))

    def run():
        model._graphs.clear()                       # capture again under the current switches
        with torch.no_grad():
            model(Ts.to(cuda), t.to(cuda), key, grasp)                       # plan + capture
            (ang, lin), _ = model(Ts.to(cuda), t.to(cuda), key, grasp)       # replay
        return ang.clone(), lin.clone()

    base = run()
    monkeypatch.setenv("DEDF_EARLY_QUERY", "0")
    late = run()
    monkeypatch.delenv("DEDF_EARLY_QUERY")
    assert torch.equal(base[0], late[0]) and torch.equal(base[1], late[1])
    ops.MLP_F16 = ops.TPACT_F16 = False
    try:
        tf32 = run()
    finally:
        ops.MLP_F16 = ops.TPACT_F16 = True
    for a, b in zip(base, tf32):
        assert torch.isfinite(a).all() and rel_err(a, b) <= 2e-5, rel_err(a, b)


def test_edge_tp_act_tc_fp16_range(cuda):
    """The fp16 operand split has fp16 RANGE: a depthwise-TP output beyond 65504 must come out non-finite (loud), never as a finite
    wrong number, and the tf32 split (DEDF_TPACT_F16=0 / f16=False) must still agree with the fp32 CUDA-core kernel there."""
    from diffusion_edf_b200 import _lib as L, layers, ops
    from oracle import so3
    G, E, n_src, n_dst = 32, 96, 50, 7
    irr = "64x0e+32x1e+16x2e"
    gen = torch.Generator().manual_seed(11)
    torch.manual_seed(11)
    pga = layers.GraphAttention(irr, irr, [32, 16, 16], 4).to(cuda)
    p = pga.packed()
    F, numel = pga.irreps_emb.dim, 15 * G
    es = torch.randint(0, n_src, (E,), generator=gen)
    ed = torch.randint(0, n_dst, (E,), generator=gen).sort().values
    row_ptr = torch.zeros(n_dst + 1, dtype=torch.long)
    row_ptr[1:] = torch.bincount(ed, minlength=n_dst).cumsum(0)
    rp = row_ptr.int().to(cuda)
    csr = ops.Csr(rp, es.int().to(cuda), ed.int().to(cuda), rp[-1:], E, n_dst, 1)
    sh = so3.spherical_harmonics(2, torch.randn(E, 3, generator=gen)).to(cuda)
    msg = (torch.randn(n_src, F, generator=gen) * 3e5).to(cuda)              # message x harmonic x weight ~ 1e5..1e6 > 65504
    w = (torch.randn(E, numel, generator=gen) / 3.0 ** 0.5).to(cuda)
    out = {}
    for name, f16 in (("f16", True), ("tf32", False)):
        logits = torch.empty(E, 4, device=cuda); v = torch.empty(E, F, device=cuda)
        ops.edge_tp_act_tc(G, msg, None, csr, sh, w, numel, p["Wtc16"] if f16 else p["Wtc"], p["b0"], p["alpha_dot"], None, logits, v, f16=f16)
        out[name] = (logits, v)
    logits2 = torch.empty(E, 4, device=cuda); v2 = torch.empty(E, F, device=cuda)
    ops.edge_tp_lin(G, L.EPI_ACT, msg, None, False, csr, sh, w, numel, p["W0"], p["W1"], p["W2"], p["b0"],
                    alpha_dot=p["alpha_dot"], edge_logit=None, logits=logits2, out=v2)
    assert torch.isfinite(v2).all()
    assert not torch.isfinite(out["f16"][1]).all(), "operands beyond the fp16 range must not produce finite values silently"
    assert torch.isfinite(out["tf32"][1]).all() and rel_err(out["tf32"][1], v2) < 2e-5
