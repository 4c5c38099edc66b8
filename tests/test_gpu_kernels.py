"""Per-kernel parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Tolerance: 1e-4 relative (max |a-b| / max |b|), the bound BASELINE.json's north_star states for fp32."""
import math
import os

import pytest
import torch

from oracle import encoders as enc
from oracle import graph as OG
from oracle import model as OM
from oracle import nn as ON
from oracle import so3
from oracle.irreps import Irreps as OIrreps
from tests.util import assert_close, random_graph, rel_err

pytestmark = pytest.mark.gpu

IRR = {32: "64x0e+32x1e+16x2e", 16: "32x0e+16x1e+8x2e"}
SH = "1x0e+1x1e+1x2e"
TOL = 1e-4


def _csr(cuda, row_ptr, edge_src, edge_dst, n_dst):
    from diffusion_edf_b200 import ops
    rp = row_ptr.to(cuda)
    return ops.Csr(rp, edge_src.to(cuda), edge_dst.to(cuda), rp[-1:], int(row_ptr[-1]), n_dst, 1)


def _random_sh(E, gen):
    v = torch.randn(E, 3, generator=gen)
    return so3.spherical_harmonics(2, v), v


# --------------------------------------------------------------------------- geometry
def test_edge_geom(cuda):
    from diffusion_edf_b200 import ops
    g = torch.Generator().manual_seed(0)
    clouds = [torch.rand(n, 3, generator=g) * 30 for n in (500, 120, 30, 8)]
    y = torch.rand(40, 3, generator=g) * 30
    y[0] = clouds[0][3]                              # zero-length edge
    radii = [5.0, 10.0, 20.0, None]
    off = [0, 500, 620, 650, 658]
    xs = torch.cat(clouds)
    csr = ops.radius_csr(xs.to(cuda), y.to(cuda), radii, src_off=off)
    length, sh, logit = ops.edge_geom(xs.to(cuda), y.to(cuda), csr, radii=radii, src_off=off, ns_cut=(0.06, 0.3), want_logit=True)
    es, ed = csr.edge_src.cpu().long(), csr.edge_dst.cpu().long()
    rp = csr.row_ptr.cpu().long()
    for s, r in enumerate(radii):
        lo, hi = int(rp[s * 40]), int(rp[(s + 1) * 40])
        par = (OM.InfiniteBipartite(SH, 0.3, 64, 100.0, fill_edge_weights=True) if r is None
               else OM.RadiusBipartite(r, SH, 64, 0.3))
        ge = par._encode_edges(xs, y, es[lo:hi], ed[lo:hi], par.fill_edge_weights if r is None else None)
        assert_close(length[lo:hi], ge.edge_length, 1e-6, f"length s{s}")
        assert_close(sh[lo:hi], ge.edge_attr, 1e-5, f"sh s{s}")
        # log(1 - soft_step(u)) cancels catastrophically as u -> 1: compare the weights, and the logits loosely
        assert (logit[lo:hi].cpu().exp() - ge.edge_logits.exp()).abs().max() < 1e-6, f"edge weight s{s}"
        assert (logit[lo:hi].cpu() - ge.edge_logits).abs().max() < 5e-3, f"logit s{s}"
    # plain variant (UNet): no cut-off, no logits
    length2, sh2, lg2 = ops.edge_geom(xs.to(cuda), y.to(cuda), csr)
    assert lg2 is None
    vec = xs[es] - y[ed]
    assert_close(sh2, so3.spherical_harmonics(2, vec), 1e-5, "plain sh")


# --------------------------------------------------------------------------- edge MLPs
@pytest.mark.parametrize("fc,numel,r", [([32, 16, 16], 240, 3.0), ([64, 32, 32], 480, 15.0), ([64, 32, 32], 240, 15.0)])
def test_edge_mlp_rbf(cuda, fc, numel, r):
    from diffusion_edf_b200 import _lib as L, layers, ops
    torch.manual_seed(1)
    E = 777
    length = torch.rand(E) * r
    length[:5] = torch.tensor([0.0, 0.001 * r, 0.01 * r, 0.1 * r, r])
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
    out = torch.empty(E, numel, device=cuda)
    d = L.MlpDesc()
    d.mode = L.MLP_IN_RBF
    d.n_edges_dev = L.ptr(n_dev, torch.int32)
    ld = length.to(cuda)
    d.length = L.ptr(ld)
    m, s, w = (p_rbf.mean.detach().reshape(-1), p_rbf.std_logit.detach().reshape(-1), p_rbf.weight_logit.detach().reshape(-1))
    d.rbf_mean, d.rbf_std_logit, d.rbf_weight_logit = L.ptr(m), L.ptr(s), L.ptr(w)
    d.rbf_cutoff, d.rbf_offset = p_rbf.cutoff, p_rbf.offset
    p_rad.fill_desc(d, 0)
    d.out = L.ptr(out)
    ops.edge_mlp(d, E)
    assert_close(out, ref, TOL, "RadialProfile(RBF)")


def test_edge_mlp_rows(cuda):
    from diffusion_edf_b200 import _lib as L, layers, ops
    torch.manual_seed(2)
    E = 301
    x = torch.randn(E, 128)
    o_rad = ON.RadialProfile([128, 128, 64, 480])
    with torch.no_grad():
        ref = o_rad(x)
    p_rad = layers.RadialProfile([128, 128, 64, 480])
    p_rad.load_state_dict(o_rad.state_dict())
    p_rad = p_rad.to(cuda)
    n_dev = torch.tensor([E], dtype=torch.int32, device=cuda)
    out = torch.empty(E, 480, device=cuda)
    xd = x.to(cuda)
    d = L.MlpDesc()
    d.mode = L.MLP_IN_ROWS
    d.n_edges_dev = L.ptr(n_dev, torch.int32)
    d.x_in = L.ptr(xd)
    p_rad.fill_desc(d, 0)
    d.out = L.ptr(out)
    ops.edge_mlp(d, E)
    assert_close(out, ref, TOL, "RadialProfile(rows)")


# --------------------------------------------------------------------------- fused TP + linear
@pytest.mark.parametrize("G", [32, 16])
def test_graph_attention_pieces(cuda, G):
    """edge_tp_lin (both epilogues) + segment_softmax_reduce against GraphAttentionMLP2's arithmetic."""
    from diffusion_edf_b200 import _lib as L, layers, ops
    gen = torch.Generator().manual_seed(10 + G)
    torch.manual_seed(10 + G)
    irr = OIrreps(IRR[G])
    F, numel = irr.dim, 15 * G
    n_src, n_dst = 90, 37
    row_ptr, es, ed = random_graph(n_src, n_dst, 9, gen)
    E = int(row_ptr[-1])
    sh, _ = _random_sh(E, gen)
    msg_src = torch.randn(n_src, F, generator=gen)
    msg_dst = torch.randn(n_dst, F, generator=gen)
    w = torch.randn(E, numel, generator=gen) / math.sqrt(3.0)
    edge_logit = -torch.rand(E, generator=gen) * 3
    oga = OM.GraphAttentionMLP2(irr, SH, irr, [32, 16, 16], 4)
    # give the zero-initialised biases values so that they are exercised
    with torch.no_grad():
        for p in oga.parameters():
            if p.abs().sum() == 0:
                p.uniform_(-0.5, 0.5)
    pga = layers.GraphAttention(IRR[G], IRR[G], [32, 16, 16], 4)
    pga.load_state_dict(oga.state_dict())
    pga = pga.to(cuda)
    with torch.no_grad():
        message = msg_src[es.long()] + msg_dst[ed.long()]
        d1 = oga.sep_act.dtp(message, sh, w)
        la = oga.sep_alpha(d1).reshape(E, 4, -1)
        v_ref = oga.sep_act.gate(oga.sep_act.lin(d1))
        la = oga.c_slrelu * ON.smooth_leaky_relu(la)
        logit_ref = torch.einsum("ehk,hk->eh", la, oga.alpha_dot.squeeze(0)) + edge_logit[:, None]
        val_ref = oga.sep_value(v_ref, edge_attr=sh, edge_scalars=None)
        logZ = OG.scatter_logsumexp(logit_ref, ed.long(), n_dst)
        alpha = torch.exp(logit_ref - logZ[ed.long()])
        attn_ref = ON.heads2vec(OG.scatter_sum(ON.vec2heads(val_ref, oga.irreps_head, 4) * alpha[..., None], ed.long(), n_dst), oga.irreps_head)
    csr = _csr(cuda, row_ptr, es, ed, n_dst)
    p = pga.packed()
    logit_ref, v_ref, val_ref = logit_ref.contiguous(), v_ref.contiguous(), val_ref.contiguous()
    logits = torch.empty(E, 4, device=cuda)
    v = torch.empty(E, F, device=cuda)
    ops.edge_tp_lin(G, L.EPI_ACT, msg_src.to(cuda), msg_dst.to(cuda), False, csr, sh.to(cuda), w.to(cuda), numel, p["W0"], p["W1"],
                    p["W2"], p["b0"], alpha_dot=p["alpha_dot"], edge_logit=edge_logit.to(cuda), logits=logits, out=v)
    assert_close(logits, logit_ref, TOL, "attention logits")
    assert_close(v, v_ref, TOL, "gated value")
    val = torch.empty(E, F, device=cuda)
    ops.edge_tp_lin(G, L.EPI_LIN, v_ref.to(cuda), None, True, csr, sh.to(cuda), p["wv"], 0, p["V0"], p["V1"], p["V2"], p["vb"], out=val)
    assert_close(val, val_ref, TOL, "sep_value")
    attn = ops.segment_softmax_reduce(csr, logit_ref.to(cuda), val_ref.to(cuda), pga.irreps_emb.m)
    assert_close(attn, attn_ref, TOL, "softmax-reduce")
    assert float(attn[0].abs().max()) == 0.0, "isolated destination must give zeros"
    # whole attend() path
    attn2 = pga.attend(msg_src.to(cuda), msg_dst.to(cuda), csr, sh.to(cuda), w.to(cuda), edge_logit.to(cuda), w_perm=False)
    assert_close(attn2, attn_ref, TOL, "attend")


@pytest.mark.parametrize("G", [32, 16])
def test_value_reduce_multisegment(cuda, G):
    """dedf_value_reduce (the value path reassociated) on a ragged multi-segment graph: destinations with no
    edges at all, with edges in one segment only, with more than one 64-edge chunk; the post-softmax factor; against the
    reference arithmetic (sep_value per edge -> scatter_logsumexp -> exp -> scatter sum, graph_attention.py:237-266) and
    against the un-reassociated kernels; bit-reproducible."""
    from diffusion_edf_b200 import _lib as L, layers, ops
    gen = torch.Generator().manual_seed(40 + G)
    torch.manual_seed(40 + G)
    irr = OIrreps(IRR[G])
    F = irr.dim
    n_dst, n_seg = 23, 3
    deg = torch.poisson(torch.full((n_seg, n_dst), 12.0), generator=gen).long()
    deg[:, 0] = 0                      # isolated destination
    deg[0, 1] = 0; deg[2, 1] = 0       # one segment only
    deg[1, 2] = 150                    # three chunks of 64
    deg[:, 3] = torch.tensor([64, 0, 64])   # exactly two full chunks
    flat = deg.reshape(-1)
    row_ptr = torch.zeros(n_seg * n_dst + 1, dtype=torch.long)
    row_ptr[1:] = flat.cumsum(0)
    E = int(row_ptr[-1])
    ed = torch.arange(n_dst).repeat(n_seg).repeat_interleave(flat)
    es = torch.randint(0, 50, (E,), generator=gen)
    sh, _ = _random_sh(E, gen)
    v = torch.randn(E, F, generator=gen)
    logit = torch.randn(E, 4, generator=gen) * 3
    post = torch.rand(E, generator=gen)
    oga = OM.GraphAttentionMLP2(irr, SH, irr, [32, 16, 16], 4)
    with torch.no_grad():
        for prm in oga.parameters():
            if prm.abs().sum() == 0:
                prm.uniform_(-0.5, 0.5)
    pga = layers.GraphAttention(IRR[G], IRR[G], [32, 16, 16], 4)
    pga.load_state_dict(oga.state_dict())
    pga = pga.to(cuda)
    p = pga.packed()
    rp = row_ptr.int().to(cuda)
    csr = ops.Csr(rp, es.int().to(cuda), ed.int().to(cuda), rp[-1:], E, n_dst, n_seg)
    for use_post in (False, True):
        with torch.no_grad():
            val_ref = oga.sep_value(v, edge_attr=sh, edge_scalars=None)
            logZ = OG.scatter_logsumexp(logit, ed, n_dst)
            alpha = torch.exp(logit - logZ[ed])
            if use_post:
                alpha = alpha * post[:, None]
            ref = ON.heads2vec(OG.scatter_sum(ON.vec2heads(val_ref, oga.irreps_head, 4) * alpha[..., None], ed, n_dst), oga.irreps_head)
        pc = post.to(cuda) if use_post else None
        out = ops.value_reduce(G, csr, v.to(cuda), sh.to(cuda), logit.to(cuda), pc, p["wv"], p["V0"], p["V1"], p["V2"], p["vb"])
        assert_close(out, ref, TOL, f"value_reduce(post={use_post})")
        assert float(out[0].abs().max()) == 0.0, "isolated destination must give zeros"
        out2 = ops.value_reduce(G, csr, v.to(cuda), sh.to(cuda), logit.to(cuda), pc, p["wv"], p["V0"], p["V1"], p["V2"], p["vb"])
        assert torch.equal(out, out2), "value_reduce must be deterministic"
        # the un-reassociated kernels (linear per edge, then the softmax-weighted sum)
        val = torch.empty(E, F, device=cuda)
        ops.edge_tp_lin(G, L.EPI_LIN, v.to(cuda), None, True, csr, sh.to(cuda), p["wv"], 0, p["V0"], p["V1"], p["V2"], p["vb"], out=val)
        if use_post:
            val = ops.row_scale(val, pc, pga.irreps_emb.m)
        old = ops.segment_softmax_reduce(csr, logit.to(cuda), val, pga.irreps_emb.m)
        assert_close(out, old, TOL, f"value_reduce vs tp_lin+softmax (post={use_post})")


@pytest.mark.parametrize("G", [32, 16])
def test_edge_tp_reduce_k1(cuda, G):
    """K1 == scatter(alpha_head(u) * o3.TensorProduct(x[src], sh, w), dst)."""
    from diffusion_edf_b200 import ops
    gen = torch.Generator().manual_seed(100 + G)
    irr = OIrreps(IRR[G])
    F, numel = irr.dim, 15 * G
    n = 200
    row_ptr, es, ed = random_graph(n, n, 13, gen)
    E = int(row_ptr[-1])
    sh, _ = _random_sh(E, gen)
    x = torch.randn(n, F, generator=gen)
    w = torch.randn(E, numel, generator=gen)
    alpha = torch.rand(E, 4, generator=gen)
    dtp = ON.DepthwiseTensorProduct(irr, OIrreps(SH), irr, internal_weights=False, bias=False)
    d = dtp(x[es.long()], sh, w)                                   # (E, 49 G)
    # alpha of the head the INPUT channel u belongs to, per output entry (depthwise: output channel == input channel)
    cols = []
    for (i1, i2, io, mode) in dtp.tp.instructions:
        pass
    head_cols = torch.empty(dtp.irreps_out.dim, dtype=torch.long)
    sl = dtp.irreps_out.slices()
    for (i1, i2, io, mode) in dtp.tp.instructions:
        m1 = irr[i1][0]
        dd = 2 * dtp.irreps_out[io][1] + 1
        hc = (torch.arange(m1) // (m1 // 4)).repeat_interleave(dd)
        head_cols[sl[io]] = hc
    ref = OG.scatter_sum(d * alpha[:, head_cols], ed.long(), n)
    out = ops.edge_tp_reduce(G, x.to(cuda), row_ptr.to(cuda), es.to(cuda), sh.to(cuda), w.to(cuda), alpha.to(cuda))
    assert_close(out, ref, TOL, "K1 (plain loads)")
    assert float(out[0].abs().max()) == 0.0
    # TMA bulk-copy pipeline: harmonics rows padded to 12 floats
    sh12 = torch.zeros(E, 12)
    sh12[:, :9] = sh
    out2 = ops.edge_tp_reduce(G, x.to(cuda), row_ptr.to(cuda), es.to(cuda), sh12.to(cuda), w.to(cuda), alpha.to(cuda))
    assert_close(out2, ref, TOL, "K1 (TMA pipeline)")
    assert float(out2[0].abs().max()) == 0.0


def test_edge_tp_reduce_k1_large_regular(cuda):
    """C4-shaped instance (degree 32, 20k nodes): the two K1 variants agree, and a linearity property holds
    (out(w1 + w2) = out(w1) + out(w2)) -- size-independent checks where the oracle would be slow."""
    from diffusion_edf_b200 import ops
    g = torch.Generator().manual_seed(7)
    N, deg = 20_000, 32
    E = N * deg
    row_ptr = (torch.arange(N + 1) * deg).int().to(cuda)
    es = torch.randint(0, N, (E,), generator=g, dtype=torch.int32).to(cuda)
    x = torch.randn(N, 240, generator=g).to(cuda)
    sh12 = torch.zeros(E, 12)
    sh12[:, :9] = so3.spherical_harmonics(2, torch.randn(E, 3, generator=g))
    sh12 = sh12.to(cuda)
    w1, w2 = torch.randn(E, 480, generator=g).to(cuda), torch.randn(E, 480, generator=g).to(cuda)
    alpha = torch.rand(E, 4, generator=g).to(cuda)
    a = ops.edge_tp_reduce(32, x, row_ptr, es, sh12, w1, alpha)
    b = ops.edge_tp_reduce(32, x, row_ptr, es, sh12[:, :9].contiguous(), w1, alpha)
    assert_close(a, b, 1e-5, "TMA vs plain")
    c = ops.edge_tp_reduce(32, x, row_ptr, es, sh12, w2, alpha)
    d = ops.edge_tp_reduce(32, x, row_ptr, es, sh12, w1 + w2, alpha)
    assert_close(d, a + c, 1e-5, "linearity in the weights")


# --------------------------------------------------------------------------- node kernels
@pytest.mark.parametrize("irr_in,irr_out,bias", [("3x0e", IRR[16], True), (IRR[16], IRR[32], True), (IRR[32], IRR[16], False),
                                                 ("112x0e+192x1e", "33x0e+32x1e", True), (IRR[32], IRR[32], True)])
def test_linear_rs(cuda, irr_in, irr_out, bias):
    from diffusion_edf_b200 import layers
    torch.manual_seed(3)
    o = ON.LinearRS(OIrreps(irr_in), OIrreps(irr_out), bias=bias)
    with torch.no_grad():
        for b in o.bias:
            b.uniform_(-1, 1)
    p = layers.LinearRS(irr_in, irr_out, bias=bias)
    p.load_state_dict(o.state_dict())
    p = p.to(cuda)
    x = torch.randn(53, OIrreps(irr_in).dim)
    with torch.no_grad():
        ref = o(x)
    assert_close(p(x.to(cuda)), ref, TOL, f"LinearRS {irr_in}->{irr_out}")


@pytest.mark.parametrize("G", [32, 16])
def test_ln_linear_gate_residual(cuda, G):
    from diffusion_edf_b200 import layers
    torch.manual_seed(4)
    irr = OIrreps(IRR[G])
    mid = ON.sort_even_first(irr * 3)[0].simplify()
    o_ln = ON.EquivariantLayerNormV2(irr)
    o_ffn = OM.FeedForwardNetwork(irr, irr, mid)
    with torch.no_grad():
        o_ln.affine_weight.uniform_(0.5, 1.5); o_ln.affine_bias.uniform_(-0.5, 0.5)
        for b in list(o_ffn.fctp_1.bias) + list(o_ffn.fctp_2.bias):
            b.uniform_(-0.5, 0.5)
    p_ln = layers.EquivariantLayerNormV2(IRR[G])
    p_ln.load_state_dict(o_ln.state_dict())
    p_ffn = layers.FeedForwardNetwork(IRR[G], IRR[G], str(mid))
    p_ffn.load_state_dict(o_ffn.state_dict())
    p_ln, p_ffn = p_ln.to(cuda), p_ffn.to(cuda)
    x = torch.randn(41, irr.dim) * torch.rand(41, 1) * 3
    with torch.no_grad():
        ref = x + o_ffn(o_ln(x))
    assert_close(p_ffn(x.to(cuda), ln=p_ln, res=x.to(cuda)), ref, TOL, "x + FFN(LN(x))")
    # LN + linear alone (ProjectIfMismatch)
    o_proj = ON.ProjectIfMismatch(OIrreps(IRR[16]), OIrreps(IRR[32]))
    p_proj = layers.ProjectIfMismatch(IRR[16], IRR[32])
    p_proj.load_state_dict(o_proj.state_dict())
    xx = torch.randn(30, 120)
    with torch.no_grad():
        ref = o_proj(xx)
    assert_close(p_proj.to(cuda)(xx.to(cuda)), ref, TOL, "ProjectIfMismatch")


@pytest.mark.parametrize("G,n", [(32, 41), (16, 41), (32, 300), (16, 2000), (32, 8)])
def test_node_chain_is_bit_identical_to_three_launches(cuda, G, n):
    """dedf_node_chain (proj -> +res -> LN -> fctp_1 -> gate -> fctp_2 -> +res in one launch) against the three dedf_node_linear
    launches it replaces: same micro-kernel and summation order, so the results must be EQUAL (above 2 x SMs nodes; the 4-node
    tiles of smaller launches split K: equal to rounding), with and without the first residual; and against the oracle (graph_attention.py:118-121, gnn_block.py:51-57,207-216)."""
    from diffusion_edf_b200 import layers, ops
    from diffusion_edf_b200.block import node_tail
    torch.manual_seed(40 + G)
    irr = OIrreps(IRR[G])
    mid = ON.sort_even_first(irr * 3)[0].simplify()
    o_ln = ON.EquivariantLayerNormV2(irr)
    o_ffn = OM.FeedForwardNetwork(irr, irr, mid)
    o_proj = ON.LinearRS(irr, irr)
    with torch.no_grad():
        o_ln.affine_weight.uniform_(0.5, 1.5); o_ln.affine_bias.uniform_(-0.5, 0.5)
        for b in list(o_ffn.fctp_1.bias) + list(o_ffn.fctp_2.bias) + list(o_proj.bias):
            b.uniform_(-0.5, 0.5)
    p_ln = layers.EquivariantLayerNormV2(IRR[G]); p_ln.load_state_dict(o_ln.state_dict())
    p_ffn = layers.FeedForwardNetwork(IRR[G], IRR[G], str(mid)); p_ffn.load_state_dict(o_ffn.state_dict())
    p_proj = layers.LinearRS(IRR[G], IRR[G]); p_proj.load_state_dict(o_proj.state_dict())
    p_ln, p_ffn, p_proj = p_ln.to(cuda), p_ffn.to(cuda), p_proj.to(cuda)
    x = torch.randn(n, irr.dim) * torch.rand(n, 1) * 3
    res = torch.randn(n, irr.dim)
    for r in (None, res):
        with torch.no_grad():
            y1 = o_proj(x) + (r if r is not None else 0.0)
            ref = y1 + o_ffn(o_ln(y1))
        rg = r.to(cuda) if r is not None else None
        assert ops.USE_NODE_CHAIN
        fused = node_tail(p_proj, p_ln, p_ffn, x.to(cuda), rg)
        ops.USE_NODE_CHAIN = False
        try:
            unfused = node_tail(p_proj, p_ln, p_ffn, x.to(cuda), rg)
        finally:
            ops.USE_NODE_CHAIN = True
        if n > 2 * 148:      # 8- / 16-node tiles: same micro-kernel and K order as dedf_node_linear
            assert torch.equal(fused, unfused), f"max diff {(fused - unfused).abs().max().item():.3e}"
        else:                # 4-node tiles split K over four lanes: same products, another association
            assert_close(fused, unfused, 2e-6, "node chain (split-K) vs three launches")
        assert_close(fused, ref, TOL, "node chain vs oracle")


def test_node_linear_pair(cuda):
    """linear_src / linear_dst of a UNet block in one launch == the two separate launches (incl. the 120 -> 240 pool block)."""
    from diffusion_edf_b200 import layers, ops
    torch.manual_seed(44)
    for irr_src, irr_dst, ns, nd in ((IRR[16], IRR[32], 400, 80), (IRR[32], IRR[32], 80, 80), (IRR[16], IRR[16], 2000, 401)):
        ls = layers.LinearRS(irr_src, irr_dst, bias=False).to(cuda)
        ld = layers.LinearRS(irr_dst, irr_dst, bias=True).to(cuda)
        with torch.no_grad():
            ld.bias[0].uniform_(-0.5, 0.5)
        xs = torch.randn(ns, ls.irreps_in.dim, device=cuda)
        xd = torch.randn(nd, ld.irreps_in.dim, device=cuda)
        (Ws, bs), (Wd, bd) = ls.packed(), ld.packed()
        ya, yb = ops.node_linear_pair(xs, ls.irreps_in.m, Ws, bs, xd, ld.irreps_in.m, Wd, bd, ld.irreps_out.m)
        assert torch.equal(ya, ls(xs)) and torch.equal(yb, ld(xd))


def test_misc_node_ops(cuda):
    from diffusion_edf_b200 import ops
    torch.manual_seed(5)
    x = torch.randn(100, 120)
    idx = torch.randint(100, (33,))
    assert torch.equal(ops.gather_rows(x.to(cuda), idx.to(cuda)).cpu(), x[idx])
    a, b = torch.randn(77, 240), torch.randn(77, 240)
    assert_close(ops.add_scale(a.to(cuda), b.to(cuda), 1 / math.sqrt(3)), (a + b) / math.sqrt(3), 1e-6, "add_scale")


# --------------------------------------------------------------------------- head kernels
def test_query_transform(cuda):
    from diffusion_edf_b200 import ops
    torch.manual_seed(6)
    irr = OIrreps(IRR[32])
    nT, nQ = 33, 5
    q = torch.randn(nT, 4)
    q[0] = torch.tensor([1.0, 0, 0, 0])                  # identity (the Euler route of the reference is singular here)
    q[1] = torch.tensor([-0.3, 0.1, 0.9, 0.2])           # w < 0, un-normalised
    q[2:] = torch.nn.functional.normalize(q[2:], dim=-1)
    Ts = torch.cat([q, torch.randn(nT, 3) * 10], -1)
    pcd = OM.FeaturedPoints(torch.randn(nQ, 3), torch.randn(nQ, irr.dim), torch.zeros(nQ, dtype=torch.long), torch.rand(nQ))
    tp = OM.TransformPcd(irr)
    # fp64 oracle: the reference's Euler-angle route loses digits in fp32 near beta -> 0 (SURVEY.md hard parts)
    ref = tp(OM.FeaturedPoints(pcd.x.double(), pcd.f.double(), pcd.b, pcd.w.double()), Ts.double())
    x, f = ops.query_transform(Ts.to(cuda), pcd.x.to(cuda), pcd.f.to(cuda), (64, 32, 16))
    assert_close(x.view(nT, nQ, 3)[1:], ref.x[1:], 1e-5, "transformed points")
    assert_close(f.view(nT, nQ, -1)[1:], ref.f[1:], 2e-5, "transformed features")
    # identity pose: D = 1 exactly in the kernel; the reference's fp64 route is accurate to ~1e-8 there
    assert_close(f.view(nT, nQ, -1)[0], pcd.f, 1e-6, "identity pose")


def test_time_embed(cuda):
    from diffusion_edf_b200 import ScoreModelHead, ops
    from diffusion_edf_b200.synthetic import model_kwargs
    torch.manual_seed(7)
    kw = model_kwargs()["score_head_kwargs"]
    tf = kw["key_tensor_field_kwargs"]
    tf.update(irreps_input="64x0e+32x1e+16x2e", use_src_point_attn=False, use_dst_point_attn=False)
    okw = dict(tf)
    ohead = OM.ScoreModelHead(1.0, [256, 128, 64], okw, "64x0e+32x1e+16x2e", 15.0, 2.5, edge_time_encoding=True, query_time_encoding=False)
    head = ScoreModelHead(1.0, [256, 128, 64], dict(tf), "64x0e+32x1e+16x2e", 15.0, 2.5, edge_time_encoding=True, query_time_encoding=False)
    head.load_state_dict(ohead.state_dict())
    head = head.to(cuda)
    t = torch.cat([torch.tensor([1e-4, 0.01, 1.0]), torch.rand(20)])
    rows = ops.time_embed(head._time_desc(), t.to(cuda))
    with torch.no_grad():
        te = ohead.time_enc(t.double()).float()          # fp64 range reduction of sin/cos(t * 1e4 * f)
        for s in range(4):
            emb = ohead.time_mlps_multiscale[s](te)
            lin = ohead.key_tensor_field.edge_scalars_pre_linears[s][0]
            ref = emb @ lin.weight[:, 64:].T + lin.bias
            assert_close(rows[s], ref, TOL, f"time rows scale {s}")


@pytest.mark.parametrize("nT,nQ", [(9, 3), (1, 1), (593, 2), (601, 1), (592, 3)])
def test_score_tp(cuda, nT, nQ):
    """Small batches run one pose per CTA, large ones (>= 4 waves) two poses per CTA and 4 (pose, query) rows per weight
    pass: odd pose counts, 1 / 2 / 3 query nodes."""
    from diffusion_edf_b200 import ops
    from diffusion_edf_b200.score_head import _ScoreTP
    from diffusion_edf_b200.irreps import Irreps
    torch.manual_seed(8)
    irr = OIrreps(IRR[32])
    pres = OIrreps("1x0e") + OIrreps("32x1e")
    o_lin = ON.SeparableFCTP(irr, irr, pres, None, use_activation=True, internal_weights=True)
    o_ang = ON.SeparableFCTP(irr, irr, pres, None, use_activation=True, internal_weights=True)
    with torch.no_grad():
        for m in (o_lin, o_ang):
            m.lin.bias[0].uniform_(-1, 1)
    p_lin, p_ang = _ScoreTP(Irreps(IRR[32]), 32), _ScoreTP(Irreps(IRR[32]), 32)
    p_lin.load_state_dict(o_lin.state_dict()); p_ang.load_state_dict(o_ang.state_dict())
    p_lin, p_ang = p_lin.to(cuda), p_ang.to(cuda)
    a, b = torch.randn(nT * nQ, 240), torch.randn(nT * nQ, 240)
    q = torch.nn.functional.normalize(torch.randn(nT, 4), dim=-1)
    Ts = torch.cat([q, torch.randn(nT, 3)], -1)
    qx, qw = torch.randn(nQ, 3) * 5, torch.rand(nQ)
    with torch.no_grad():
        lin = o_lin(a, b, edge_scalars=None)[..., 1:].view(nT, nQ, 32, 3).mean(-2)
        ang = o_ang(a, b, edge_scalars=None)[..., 1:].view(nT, nQ, 32, 3).mean(-2)
        qinv = enc.quaternion_invert(q.unsqueeze(-2))
        lin, ang = enc.quaternion_apply(qinv, lin), enc.quaternion_apply(qinv, ang)
        orb = torch.cross(qx.unsqueeze(0) / 15.0, lin, dim=-1)
        lin_ref = torch.einsum("q,tqi->ti", qw, lin)
        ang_ref = torch.einsum("q,tqi->ti", qw, orb) + torch.einsum("q,tqi->ti", qw, ang)
    Wd = [p_lin.packed_dtp(), p_ang.packed_dtp()]
    (l0, l1, _), lb = p_lin.lin.packed()
    (a0, a1, _), ab = p_ang.lin.packed()
    ang_g, lin_g = ops.score_tp(Ts.to(cuda), a.to(cuda), b.to(cuda), qx.to(cuda), qw.to(cuda), (64, 32, 16), Wd, [l0, a0], [l1, a1], [lb, ab], 32, 15.0)
    assert_close(lin_g, lin_ref, TOL, "lin score")
    assert_close(ang_g, ang_ref, TOL, "ang score")


def test_pose_update(cuda):
    from diffusion_edf_b200 import ops
    torch.manual_seed(9)
    n = 50
    q = torch.nn.functional.normalize(torch.randn(n, 4, dtype=torch.float64), dim=-1)
    T = torch.cat([q, torch.randn(n, 3, dtype=torch.float64)], -1)
    ang, lin = torch.randn(n, 3), torch.randn(n, 3)
    z = torch.randn(n, 6, dtype=torch.float64)
    t, am, lm, temp = 0.37, 2.5, 15.0, 0.8
    a_ang, a_lin = am ** 2 * t ** 0.5 * 0.04, lm ** 2 * t ** 0.5 * 0.04
    # reference arithmetic (score_model_base.py:178-193)
    s_a = ang.double() / (am * math.sqrt(t)); s_l = lin.double() / (lm * math.sqrt(t))
    ad = (a_ang / 2) * s_a + math.sqrt(temp * a_ang) * z[:, :3]
    ld = (a_lin / 2) * s_l + math.sqrt(temp * a_lin) * z[:, 3:]
    qi = torch.tensor([[1, 2, 3], [0, 3, 2], [3, 0, 1], [2, 1, 0]])
    qf = torch.tensor([[-0.5, -0.5, -0.5], [0.5, -0.5, 0.5], [0.5, 0.5, -0.5], [-0.5, 0.5, 0.5]], dtype=torch.float64)
    Lm = T[..., qi] * qf
    dq = torch.einsum("...ij,...j->...i", Lm, ad)
    dx = enc.quaternion_apply(T[:, :4], ld)
    ref = torch.cat([enc.normalize_quaternion(T[:, :4] + dq), T[:, 4:] + dx], -1)
    Td = T.to(cuda).clone()
    traj = torch.empty(n, 7, dtype=torch.float64, device=cuda)
    T32 = torch.empty(n, 7, dtype=torch.float32, device=cuda)
    ops.pose_update(Td, ang.to(cuda), lin.to(cuda), z.to(cuda), 0, 0, t, am, lm, a_ang, a_lin, temp, traj, T32)
    assert (Td.cpu() - ref).abs().max() < 1e-12
    assert torch.equal(traj, Td) and torch.equal(T32, Td.float())
    # Philox path: unit quaternions, finite, different poses get different noise, deterministic in (seed, offset)
    T1, T2 = T.to(cuda).clone(), T.to(cuda).clone()
    ops.pose_update(T1, ang.to(cuda), lin.to(cuda), None, 123, 5, t, am, lm, a_ang, a_lin, temp, None, None)
    ops.pose_update(T2, ang.to(cuda), lin.to(cuda), None, 123, 5, t, am, lm, a_ang, a_lin, temp, None, None)
    assert torch.equal(T1, T2) and torch.isfinite(T1).all()
    assert (T1[:, :4].norm(dim=-1) - 1).abs().max() < 1e-12
    assert (T1 - Td).abs().max() > 1e-3


def test_pose_update_philox_steps_are_disjoint(cuda):
    """The Philox noise of different steps must not overlap (round-1 bug: offset = step counted single 32-bit outputs while a
    step consumes 12, so step s+4 re-read step s's normals).  With zero scores, identity rotation and temperature 1 the linear
    displacement IS sqrt(alpha_lin) z[3:6] and the quaternion increment is 0.5 sqrt(alpha_ang) z[0:3]: recover z per step."""
    from diffusion_edf_b200 import ops
    n, steps = 512, 12
    zero = torch.zeros(n, 3, device=cuda)
    a_ang = a_lin = 1e-6          # tiny: the first-order quaternion update is then exact to 1e-12
    zs = []
    for s in range(steps):
        T = torch.zeros(n, 7, dtype=torch.float64, device=cuda); T[:, 0] = 1.0
        ops.pose_update(T, zero, zero, None, 77, s, 0.5, 2.5, 15.0, a_ang, a_lin, 1.0, None, None)
        q = T[:, :4] / T[:, :1]
        z = torch.cat([2.0 * q[:, 1:] / math.sqrt(a_ang), T[:, 4:] / math.sqrt(a_lin)], dim=-1)
        zs.append(z.cpu())
    Z = torch.stack(zs)                                   # (steps, n, 6)
    assert torch.isfinite(Z).all()
    # no value of one step reappears in another step of the same pose (the old bug reproduced z[2:4] of step s as z[0:2] of s+4)
    for s in range(steps):
        for d in (1, 2, 3, 4, 8):
            if s + d < steps:
                a, b = Z[s], Z[s + d]
                same = (a[:, :, None] - b[:, None, :]).abs() < 1e-10      # z is recovered to ~1e-13; chance coincidence ~1e-5
                assert not same.any(), f"steps {s} and {s + d} share a normal draw"
    # standard normal, uncorrelated across steps and components
    flat = Z.permute(1, 0, 2).reshape(n, -1)             # (n, steps*6) samples of a 72-dim vector
    assert abs(float(flat.mean())) < 0.02 and abs(float(flat.std()) - 1.0) < 0.02
    C = torch.corrcoef(flat.T)
    off = C - torch.eye(C.shape[0], dtype=C.dtype)
    assert float(off.abs().max()) < 0.25, float(off.abs().max())      # 512 samples: |r| ~ 0.044 sigma, 0.25 = 5.6 sigma


# --------------------------------------------------------------------------- fused head step
@pytest.mark.parametrize("nT,nQ,batched", [(5, 2, False), (37, 3, True), (128, 2, False), (300, 1, False)])
def test_head_front_is_bit_identical_to_the_unfused_kernels(cuda, nT, nQ, batched):
    """dedf_head_front (pose transform + multi-scale radius search + CSR + edge geometry, one launch) against
    dedf_query_transform + dedf_radius_count/_fill + dedf_edge_geom: identical indices and bit-identical floats; eager (exact
    sizing) and capacity mode (clamp + overflow flag)."""
    from diffusion_edf_b200 import ops
    g = torch.Generator().manual_seed(500 + nT)
    sizes = [700, 150, 40, 9]
    radii = [5.0, 10.0, 20.0, None]
    xs = torch.cat([(torch.rand(n, 3, generator=g) - 0.5) * 40 for n in sizes]).to(cuda)
    off = [0]
    for n in sizes:
        off.append(off[-1] + n)
    b_src = (torch.cat([torch.arange(n) % 2 for n in sizes]) if batched else torch.zeros(sum(sizes), dtype=torch.long)).to(cuda)
    qx = (torch.randn(nQ, 3, generator=g) * 3).to(cuda)
    b_q = (torch.arange(nQ) % 2 if batched else torch.zeros(nQ, dtype=torch.long)).to(cuda)
    q = torch.nn.functional.normalize(torch.randn(nT, 4, generator=g), dim=-1)
    Ts = torch.cat([q, (torch.rand(nT, 3, generator=g) - 0.5) * 40], -1).to(cuda)
    # un-fused
    x_ref, _ = ops.query_transform(Ts, qx, torch.zeros(nQ, 240, device=cuda), (64, 32, 16))
    bq = b_q.unsqueeze(0).expand(nT, -1).reshape(-1).contiguous()
    g0 = ops.radius_csr(xs, x_ref, radii, src_off=off, b_src=b_src, b_dst=bq, max_num_neighbors=1000)
    l0, sh0, lg0 = ops.edge_geom(xs, x_ref, g0, radii=radii, src_off=off, ns_cut=(0.06, 0.3), want_logit=True)
    # fused, exact sizing
    g1, l1, sh1, lg1, x1 = ops.head_front(Ts, qx, b_q, xs, b_src, off, radii, (0.06, 0.3))
    assert g1.n_edges == g0.n_edges and g0.n_edges > 0
    assert torch.equal(x1, x_ref)
    assert torch.equal(g1.row_ptr, g0.row_ptr) and torch.equal(g1.edge_src, g0.edge_src) and torch.equal(g1.edge_dst, g0.edge_dst)
    E = g0.n_edges
    assert torch.equal(l1[:E], l0[:E]) and torch.equal(sh1[:E], sh0[:E]) and torch.equal(lg1[:E], lg0[:E])
    assert int(g1.n_edges_dev) == E
    # capacity mode: roomy -> same graph, flag clear; tight -> clamped CSR, flag raised
    ovf = torch.zeros(1, dtype=torch.int32, device=cuda)
    g2, l2, sh2, lg2, _ = ops.head_front(Ts, qx, b_q, xs, b_src, off, radii, (0.06, 0.3), capacity=E + 100, overflow=ovf)
    assert int(ovf) == 0 and int(g2.n_edges_dev) == E and torch.equal(g2.row_ptr, g0.row_ptr)
    assert torch.equal(g2.edge_src[:E], g0.edge_src) and torch.equal(sh2[:E], sh0[:E])
    g3, *_ = ops.head_front(Ts, qx, b_q, xs, b_src, off, radii, (0.06, 0.3), capacity=max(1, E // 2), overflow=ovf)
    assert int(ovf) == 1 and int(g3.n_edges_dev) == max(1, E // 2) and int(g3.row_ptr.max()) == max(1, E // 2)
    assert torch.equal(g3.edge_src[:E // 2], g0.edge_src[:E // 2])


def test_head_front_all_pairs_scale_is_not_capped(cuda):
    """InfiniteBipartite builds the full meshgrid: max_num_neighbors does not apply to it (graph_parser.py:274-278); both the
    fused front and dedf_radius_count/_fill must keep every source of an r = None scale."""
    from diffusion_edf_b200 import ops
    g = torch.Generator().manual_seed(7)
    xs = (torch.rand(50, 3, generator=g) * 4).to(cuda)
    qx = torch.zeros(1, 3, device=cuda)
    Ts = torch.tensor([[1.0, 0, 0, 0, 1, 1, 1], [1.0, 0, 0, 0, 2, 2, 2]], device=cuda)
    g1, *_ = ops.head_front(Ts, qx, None, xs, None, [0, 50], [None], (0.0, -1.0), max_num_neighbors=8)
    assert g1.n_edges == 100
    x_dst = Ts[:, 4:].contiguous()
    g0 = ops.radius_csr(xs, x_dst, [None], src_off=[0, 50], max_num_neighbors=8)
    assert g0.n_edges == 100 and torch.equal(g0.edge_src, g1.edge_src)
    # a finite radius IS capped
    g2 = ops.radius_csr(xs, x_dst, [100.0], src_off=[0, 50], max_num_neighbors=8)
    assert g2.n_edges == 16


@pytest.mark.parametrize("nT,nQ", [(7, 2), (130, 2), (300, 3)])
def test_score_tp_step_matches_unfused(cuda, nT, nQ):
    """dedf_score_tp_step: D(q) psi applied inside == dedf_query_transform + dedf_score_tp (bit-identical), and the fused float64
    Langevin step == dedf_pose_update on those scores (bit-identical), incl. the step counter, trajectory row and fp32 pose copy."""
    from diffusion_edf_b200 import ScoreModelHead, ops
    from diffusion_edf_b200.denoise import StepState
    from diffusion_edf_b200.synthetic import model_kwargs
    torch.manual_seed(60 + nT)
    kw = model_kwargs()["score_head_kwargs"]
    tf = kw["key_tensor_field_kwargs"]
    tf.update(irreps_input="64x0e+32x1e+16x2e", use_src_point_attn=False, use_dst_point_attn=False)
    head = ScoreModelHead(1.0, [256, 128, 64], dict(tf), "64x0e+32x1e+16x2e", 15.0, 2.5, edge_time_encoding=True, query_time_encoding=False).to(cuda)
    Wd, Wl0, Wl1, bl = head._tp_packed()
    q = torch.nn.functional.normalize(torch.randn(nT, 4), dim=-1)
    T64 = torch.cat([q, torch.randn(nT, 3) * 5], -1).double().to(cuda)
    Ts = T64.float()
    qx, qf, qw = torch.randn(nQ, 3, device=cuda), torch.randn(nQ, 240, device=cuda), torch.rand(nQ, device=cuda)
    key_f = torch.randn(nT * nQ, 240, device=cuda)
    _, fq = ops.query_transform(Ts, qx, qf, (64, 32, 16))
    ang0, lin0 = ops.score_tp(Ts, fq, key_f, qx, qw, (64, 32, 16), Wd, Wl0, Wl1, bl, 32, 15.0)
    ang1, lin1 = ops.score_tp_step(Ts, qf, key_f, qx, qw, (64, 32, 16), Wd, Wl0, Wl1, bl, 32, 15.0)
    assert torch.equal(ang0, ang1) and torch.equal(lin0, lin1)
    # fused step, Philox noise, step index 3 of a 5-step schedule
    n_steps, step = 5, 3
    sched = torch.rand(n_steps, 4, dtype=torch.float64, device=cuda) + 0.1
    st = StepState(T64.clone(), sched, torch.full((1,), step, dtype=torch.int32, device=cuda), None,
                   torch.full((1,), 99, dtype=torch.int64, device=cuda), torch.zeros(n_steps + 2, nT, 7, dtype=torch.float64, device=cuda),
                   torch.zeros(1, dtype=torch.int32, device=cuda), torch.zeros(4, n_steps, 128, device=cuda), torch.zeros(4, 1, 128, device=cuda), 2.5, 15.0)
    T32 = Ts.clone()
    ang2, lin2 = ops.score_tp_step(T32, qf, key_f, qx, qw, (64, 32, 16), Wd, Wl0, Wl1, bl, 32, 15.0, state=st)
    assert torch.equal(ang2, ang0) and torch.equal(lin2, lin0)
    ref = T64.clone()
    row = sched[step].tolist()
    ops.pose_update(ref, ang0, lin0, None, 99, step, row[0], 2.5, 15.0, row[1], row[2], row[3], None, None)
    assert torch.equal(st.T64, ref), float((st.T64 - ref).abs().max())
    assert torch.equal(st.traj[step + 1], ref) and torch.equal(T32, ref.float())
    assert int(st.counter) == step + 1 and int(st.ticket) == 0
    # injected noise rows
    noise = torch.randn(n_steps, nT, 6, dtype=torch.float64, device=cuda)
    st2 = st._replace(T64=T64.clone(), counter=torch.full((1,), step, dtype=torch.int32, device=cuda), noise=noise)
    T32 = Ts.clone()
    ops.score_tp_step(T32, qf, key_f, qx, qw, (64, 32, 16), Wd, Wl0, Wl1, bl, 32, 15.0, state=st2)
    ref = T64.clone()
    ops.pose_update(ref, ang0, lin0, noise[step].contiguous(), 0, 0, row[0], 2.5, 15.0, row[1], row[2], row[3], None, None)
    assert torch.equal(st2.T64, ref)
