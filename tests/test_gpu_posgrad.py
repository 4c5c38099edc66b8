"""Adjoints w.r.t. the query coordinates (csrc/train.cu "position gradients", the kernels behind EbmScoreModelHead.forward,
/root/reference/diffusion_edf/score_head_ebm.py:192-222) against torch autograd through the CPU oracle in float64.  The model-level
check (score of the energy-based head vs autograd through the oracle / the reference's own source) is in tests/test_gpu_model.py
::test_ebm_critic_energy and tests/test_gpu_zz_reference_golden.py[ebm]."""
import math

import pytest
import torch

from tests.util import assert_close

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _graph(n_src, n_dst, deg, gen):
    es = torch.randint(0, n_src, (n_dst * deg,), generator=gen)
    ed = torch.arange(n_dst).repeat_interleave(deg)
    return es, ed


def test_edge_geom_bwd(cuda):
    """(g_len, g_sh, g_logit) -> dx_dst of graph_parser.py:146-224 (length, normalised l <= 2 harmonics with the non-scalar min-cut,
    log of the soft edge cut-off), two scales: a radius scale (r = 2) and the all-pairs scale (logit 0)."""
    from diffusion_edf_b200 import _lib as L
    from diffusion_edf_b200 import ops
    from diffusion_edf_b200._lib import ptr
    from oracle import encoders as enc
    from oracle import so3
    gen = torch.Generator().manual_seed(0)
    n0, n1, nd, deg = 40, 10, 24, 6
    # coordinates rounded to fp32 first: the float64 reference and the kernel see the same inputs (the edge logit's derivative
    # 1 / cut is ill-conditioned near the cut-off radius)
    xs = (torch.rand(n0 + n1, 3, generator=gen, dtype=torch.float64) * 2 - 1).float().double()
    xd = (torch.rand(nd, 3, generator=gen, dtype=torch.float64) * 2 - 1).float().double().requires_grad_(True)
    es0, ed0 = _graph(n0, nd, deg, gen)
    es1, ed1 = _graph(n1, nd, 3, gen)
    es, ed = torch.cat([es0, es1 + n0]), torch.cat([ed0, ed1])
    r, ns = 2.0, 0.6
    # edges whose length lies just inside the cut-off radius are dropped: d logit / d len = -cut' / cut with cut = 1 - soft_step -> 0 is
    # ill-conditioned there in fp32 (the kernel and the reference both clamp at cut = 1e-12; the model-level tests cover that regime)
    ln0 = (xs[es] - xd[ed]).norm(dim=1).detach()
    keep = ~((ln0 > 0.93 * r) & (ln0 < 1.001 * r) & (es < n0))
    es, ed = es[keep], ed[keep]
    E = len(es)
    vec = xs[es] - xd[ed]
    ln = vec.norm(dim=1)
    cut_ns = enc.soft_square_cutoff_2(ln, (0.2 * ns, 1.0 * ns, None, None))
    sh = so3.spherical_harmonics(2, vec, normalize=True)
    sh = torch.cat([sh[:, :1], sh[:, 1:] * cut_ns[:, None]], dim=1)
    cut = enc.soft_square_cutoff_2(ln, (None, None, 0.8 * r, 1.0 * r))
    logit = torch.where(es < n0, torch.log(torch.clamp(cut, min=1e-12)), torch.zeros_like(ln))
    assert float((ln > 0.8 * r).double().mean()) > 0.02 and float((ln < ns).double().mean()) > 0.02     # both cut-offs are exercised
    g_len, g_sh, g_logit = (torch.randn(E, generator=gen, dtype=torch.float64), torch.randn(E, 9, generator=gen, dtype=torch.float64),
                            torch.randn(E, generator=gen, dtype=torch.float64))
    ((ln * g_len).sum() + (sh * g_sh).sum() + (logit * g_logit).sum()).backward()
    f = lambda t: t.detach().to(torch.float32).to(cuda).contiguous()       # noqa: E731
    dx = torch.zeros(nd, 3, dtype=torch.float32, device=cuda)
    dev = [f(xs), f(xd), es.int().to(cuda), ed.int().to(cuda), f(g_len), f(g_sh), f(g_logit)]       # kept alive across the launch
    ops._call("dedf_edge_geom_bwd", ptr(dev[0]), ptr(dev[1]), ptr(dev[2], torch.int32), ptr(dev[3], torch.int32), E, 2,
              L.int_array([0, n0, n0 + n1]), L.float_array([r, -1.0]), 0.2 * ns, 1.0 * ns, ptr(dev[4]), ptr(dev[5]), ptr(dev[6]),
              ptr(dx), L.stream())
    assert_close(dx, xd.grad, TOL, "dx_dst")


@pytest.mark.parametrize("mode", [0, 1])
def test_rbf_and_sinusoid_bwd_len(cuda, mode):
    from diffusion_edf_b200 import ops
    from diffusion_edf_b200._lib import ptr, stream
    from oracle import encoders as enc
    gen = torch.Generator().manual_seed(1)
    E, K = 300, 16
    if mode == 0:
        mod = enc.GaussianRadialBasis(dim=K, max_val=2.0).double()
        offset, inv_span = 0.0, 0.5
        pm = mod.param_module
    else:
        mod = enc.GaussianRadialBasisLayerFiniteCutoff(num_basis=K, cutoff=1.98, offset=0.1).double()
        offset, inv_span = 0.1, 1.0 / (1.98 - 0.1)
        pm = mod
    ln = (torch.rand(E, generator=gen, dtype=torch.float64) * 1.9 + 0.05).requires_grad_(True)
    g = torch.randn(E, K, generator=gen, dtype=torch.float64)
    (mod(ln) * g).sum().backward()
    f = lambda t: t.detach().to(torch.float32).reshape(-1).to(cuda).contiguous()       # noqa: E731
    dlen = torch.empty(E, dtype=torch.float32, device=cuda)
    dev = [f(ln), f(pm.mean), f(pm.std_logit), f(pm.weight_logit), g.float().to(cuda).contiguous()]  # kept alive across the launch
    ops._call("dedf_rbf_bwd_len", ptr(dev[0]), E, K, ptr(dev[1]), ptr(dev[2]), ptr(dev[3]), offset, inv_span, mode, ptr(dev[4]), ptr(dlen), stream())
    assert_close(dlen, ln.grad, TOL, f"rbf dlen mode {mode}")
    if mode == 0:
        dim, max_r = 16, 10.0
        sin = enc.SinusoidalPositionEmbeddings(dim=dim, max_val=max_r, n=1000.0)
        x = (torch.rand(E, generator=gen, dtype=torch.float64) * 8).requires_grad_(True)
        gs = torch.randn(E, dim, generator=gen, dtype=torch.float64)
        (sin(x) * gs).sum().backward()
        half = dim // 2
        freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(1000.0) / (half - 1))).to(cuda)
        dx = torch.empty(E, dtype=torch.float32, device=cuda)
        dev = [f(x), gs.float().to(cuda).contiguous()]
        ops._call("dedf_sinusoid_bwd", ptr(dev[0]), E, dim, ptr(freq), 1000.0 / max_r, ptr(dev[1]), ptr(dx), stream())
        assert_close(dx, x.grad, TOL, "sinusoid dx")


@pytest.mark.parametrize("G,per_edge", [(32, True), (16, True), (32, False)])
def test_dtp_bwd_sh(cuda, G, per_edge):
    """The depthwise tensor product is linear in the harmonics: dsh[e, j] = <g[e], dtp(x[e], e_j, w[e])> with the (oracle-tested)
    forward kernel evaluated on the nine unit vectors."""
    from diffusion_edf_b200 import ops
    from diffusion_edf_b200._lib import ptr, stream
    gen = torch.Generator().manual_seed(2)
    E, F, NW, FO = 77, 7.5 * G, 15 * G, 49 * G
    F = int(F)
    x = torch.randn(E, F, generator=gen).to(cuda)
    w = (torch.randn(E, NW, generator=gen) if per_edge else torch.randn(NW, generator=gen)).to(cuda)
    g = torch.randn(E, FO, generator=gen).to(cuda)
    stride = NW if per_edge else 0
    ref = torch.empty(E, 9, dtype=torch.float64)
    for j in range(9):
        sh = torch.zeros(E, 9, device=cuda); sh[:, j] = 1.0
        out = torch.empty(E, FO, device=cuda)
        ops._call("dedf_dtp_fwd", G, ptr(x), ptr(sh), ptr(w), stride, E, ptr(out), stream())
        ref[:, j] = (out.double() * g.double()).sum(1).cpu()
    dsh = torch.empty(E, 9, device=cuda)
    ops._call("dedf_dtp_bwd_sh", G, ptr(x), ptr(w), stride, ptr(g), E, ptr(dsh), stream())
    assert_close(dsh, ref, TOL, "dsh")


def test_ebm_pose_grad_matches_autograd(cuda):
    """dedf_ebm_pose_grad (closed-form pull-back: generators of the l = 1 / l = 2 representations) against torch autograd through the
    oracle's TransformPcd (quaternion_to_matrix + YXY-Euler Wigner matrices, wigner.py:257-283) contracted with L(q) as
    score_head_ebm.py:211-214 does."""
    from diffusion_edf_b200 import _lib as L
    from diffusion_edf_b200 import ops
    from diffusion_edf_b200._lib import ptr
    from oracle import encoders as enc
    from oracle import model as OM
    from oracle.irreps import Irreps
    gen = torch.Generator().manual_seed(3)
    irr = Irreps("8x0e+4x1e+2x2e")
    nT, nQ, F = 5, 7, irr.dim
    q = torch.randn(nT, 4, generator=gen, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)
    T = torch.cat([q, torch.randn(nT, 3, generator=gen, dtype=torch.float64)], dim=1).requires_grad_(True)
    qx = torch.randn(nQ, 3, generator=gen, dtype=torch.float64)
    qf = torch.randn(nQ, F, generator=gen, dtype=torch.float64)
    g_x = torch.randn(nT, nQ, 3, generator=gen, dtype=torch.float64)
    g_f = torch.randn(nT, nQ, F, generator=gen, dtype=torch.float64)
    pcd = OM.FeaturedPoints(qx, qf, torch.zeros(nQ, dtype=torch.long))
    out = OM.TransformPcd(irr).double()(pcd, T)
    ((out.x * g_x).sum() + (out.f * g_f).sum()).backward()
    q_indices = torch.tensor([[1, 2, 3], [0, 3, 2], [3, 0, 1], [2, 1, 0]])
    q_factor = torch.tensor([[-0.5, -0.5, -0.5], [0.5, -0.5, 0.5], [0.5, 0.5, -0.5], [-0.5, 0.5, 0.5]], dtype=torch.float64)
    Lq = T.detach()[:, q_indices] * q_factor
    ang_mult, lin_mult = 1.7, 0.3
    ang_ref = torch.einsum("tia,ti->ta", Lq, T.grad[:, :4]) * ang_mult
    lin_ref = enc.quaternion_apply(enc.quaternion_invert(T.detach()[:, :4]), T.grad[:, 4:]) * lin_mult
    f = lambda t: t.detach().to(torch.float32).to(cuda).contiguous()       # noqa: E731
    ang = torch.empty(nT, 3, device=cuda)
    lin = torch.empty(nT, 3, device=cuda)
    dev = [f(T), f(qx), f(qf), f(g_x.reshape(-1, 3)), f(g_f.reshape(-1, F))]
    ops._call("dedf_ebm_pose_grad", ptr(dev[0]), nT, nQ, L.int_array([8, 4, 2]), ptr(dev[1]), ptr(dev[2]), ptr(dev[3]), ptr(dev[4]),
              ang_mult, lin_mult, ptr(ang), ptr(lin), L.stream())
    assert_close(ang, ang_ref, TOL, "ang")
    assert_close(lin, lin_ref, TOL, "lin")
