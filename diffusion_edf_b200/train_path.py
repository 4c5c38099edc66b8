"""Differentiable (training) forward of the score network, composed from the un-fused primitives of autograd_ops.py.

Same modules / parameters / state_dict as the inference path; only the execution differs: the inference kernels fuse
gather + tensor product + linear and never materialise the (E, 49 G) tensor-product outputs, which a backward pass needs
for the weight gradients.  Mirrors the reference's arithmetic order (not the inference path's exact reassociations):
  gnn_block.py:164-218 / block.py:141-174 (EquiformerBlock), graph_attention.py:84-122, :218-273 (GraphAttentionMLP(2)),
  multiscale_tensor_field.py:192-260, graph_parser.py:146-224, unet_feature_extractor.py:260-417, score_head.py:142-211.
Train-mode dropout (alpha_drop on the attention weights, proj_drop = EquivariantDropout after GraphAttention.proj AND after
the FeedForwardNetwork's fctp_2, gnn_block.py:44-57) is applied when the module is in train() mode, with Philox masks; the
parity tests compare against the oracle in eval mode.  drop_path_rate > 0 (GraphDropPath) is rejected by the constructors.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch

from . import autograd_ops as A
from . import ops
from .gnn_data import FeaturedPoints
from .irreps import gate_pre
from .layers import EquivariantLayerNormV2, GraphAttention, LinearRS, ProjectIfMismatch, RadialProfile


# ------------------------------------------------------------------------------------------------ building blocks
def linear_rs(mod: LinearRS, x: torch.Tensor) -> torch.Tensor:
    bias = mod.bias[0] if len(mod.bias) else None
    return A.LinearFn.apply(x, mod.tp.weight, bias, mod.irreps_in.m, mod.irreps_out.m)


def layer_norm(mod: EquivariantLayerNormV2, x: torch.Tensor) -> torch.Tensor:
    return A.LayerNormFn.apply(x, mod.affine_weight, mod.affine_bias, mod.irreps.m, mod.eps)


def nn_linear(lin: torch.nn.Linear, x: torch.Tensor, extra_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    k, n = lin.in_features, lin.out_features
    w = lin.weight.t().contiguous().view(-1)          # (in x out) row-major = the LinearRS block layout (autograd-tracked)
    bias = lin.bias if lin.bias is not None else extra_bias
    assert lin.bias is None or extra_bias is None
    return A.LinearFn.apply(x, w, bias, (k, 0, 0), (n, 0, 0))


def radial_profile(mod: RadialProfile, x: torch.Tensor) -> torch.Tensor:
    mods = list(mod.net)
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, torch.nn.Linear):
            last = i == len(mods) - 1
            x = nn_linear(m, x, extra_bias=mod.offset if last else None)       # offset rides as the bias of the bias-free last layer
        elif isinstance(m, torch.nn.LayerNorm):
            x = A.LayerNormFn.apply(x, m.weight, m.bias, (m.normalized_shape[0], 0, 0), m.eps)
        elif isinstance(m, torch.nn.SiLU):
            x = A.SiluFn.apply(x)
        else:
            raise NotImplementedError(type(m))
        i += 1
    return x


def project_if_mismatch(mod: ProjectIfMismatch, x: torch.Tensor) -> torch.Tensor:
    if mod.is_identity:
        return x
    if isinstance(mod.layernorm, EquivariantLayerNormV2):
        x = layer_norm(mod.layernorm, x)
    return linear_rs(mod.skip, x)


def ffn(mod, x: torch.Tensor, proj_drop: float = 0.0) -> torch.Tensor:
    """FeedForwardNetwork.forward (gnn_block.py:51-57 / block.py:51-57): fctp_1 (+gate) -> fctp_2 -> EquivariantDropout(proj_drop)
    on the output irreps in train mode."""
    h = linear_rs(mod.fctp_1, x)
    h = A.GateFn.apply(h, mod.fctp_1.irreps_out.m)
    out = linear_rs(mod.fctp_2, h)
    if proj_drop > 0.0:
        irr = mod.fctp_2.irreps_out
        out = A.GroupScaleFn.apply(out, A.dropout_mask((out.shape[0], irr.num_irreps), proj_drop, out.device), irr.m, 1)
    return out


def graph_attention(ga: GraphAttention, msg_src: torch.Tensor, msg_dst: Optional[torch.Tensor], g: ops.Csr, sh: torch.Tensor,
                    w: torch.Tensor, edge_logit: Optional[torch.Tensor], drop=(0.0, 0.0),
                    src_weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """-> proj(sum_e softmax_e value_e) incl. the train-mode dropouts ``drop`` = (alpha_drop, proj_drop)."""
    G = ga.irreps_emb.m[1]
    E = g.n_edges
    es, ed = g.edge_src[:E].contiguous(), g.edge_dst[:E].contiguous()
    message = A.GatherFn.apply(msg_src, es)
    if msg_dst is not None:
        message = A.AddScaleFn.apply(message, A.GatherFn.apply(msg_dst, ed), 1.0)
    m = A.dtp(ga.sep_act, message, sh, w)                                                  # (E, 49 G) for the 2G:G:G/2 family
    d_out = ga.sep_act.irreps_dtp_out
    alpha_pre = linear_rs(ga.sep_alpha, m[:, :d_out.m[0]].contiguous())
    logits = A.AlphaFn.apply(alpha_pre, ga.alpha_dot, edge_logit)
    v = A.GateFn.apply(linear_rs(ga.sep_act.lin, m), ga.sep_act.lin.irreps_out.m)
    m2 = A.dtp(ga.sep_value, v, sh, ga.sep_value.dtp.tp.weight)
    val = linear_rs(ga.sep_value.lin, m2)
    if src_weight is not None:      # source-point attention: alpha_e *= w[src_e] after the softmax == scaling the value rows
        val = A.RowScaleFn.apply(val, A.GatherFn.apply(src_weight.reshape(-1, 1), es), ga.irreps_emb.m)
    if drop[0] > 0.0:       # nn.Dropout on alpha (E, heads): sum_e (alpha_e m_e) v_e == sum_e alpha_e (m_e v_e)
        val = A.GroupScaleFn.apply(val, A.dropout_mask((E, 4), drop[0], val.device), ga.irreps_emb.m, 0)
    out = linear_rs(ga.proj, A.SoftmaxReduceFn.apply(logits, val, g, ga.irreps_emb.m))
    if drop[1] > 0.0:       # EquivariantDropout: one Bernoulli per (node, irrep channel)
        irr = ga.proj.irreps_out
        out = A.GroupScaleFn.apply(out, A.dropout_mask((out.shape[0], irr.num_irreps), drop[1], out.device), irr.m, 1)
    return out


def unet_block(blk, f_src, f_dst, geom, radial, drop=(0.0, 0.0)) -> torch.Tensor:
    """block.EquiformerBlock (UNet): norm_1_* are computed-and-discarded in the reference (block.py:149-153)."""
    msg_src = linear_rs(blk.linear_src, f_src)
    msg_dst = linear_rs(blk.linear_dst, f_dst)
    E = geom.g.n_edges
    emb = A.RbfFn.apply(geom.length[:E], radial.mean, radial.std_logit, radial.weight_logit, radial.offset,
                        1.0 / (radial.cutoff - radial.offset), 1)
    w = radial_profile(blk.ga.sep_act.dtp_rad, emb)
    attn = graph_attention(blk.ga, msg_src, msg_dst, geom.g, geom.sh[:E], w, None, drop)
    out = A.AddScaleFn.apply(attn, f_dst, 1.0)
    return A.AddScaleFn.apply(ffn(blk.ffn, layer_norm(blk.norm_2, out), drop[1]), out, 1.0)


# ------------------------------------------------------------------------------------------------ key encoder
def unet_forward(net, pcd: FeaturedPoints) -> List[FeaturedPoints]:
    x, b = pcd.x.contiguous(), pcd.b.contiguous()
    f = linear_rs(net.input_emb, pcd.f.contiguous())
    outs, graphs = [(f, x, b)], []
    scale_outs = []
    geom = None

    drop = (net.alpha_drop, net.proj_drop) if net.training else (0.0, 0.0)

    def run(layer, f_src, f_dst, gm):
        return unet_block(layer["gnn"], f_src, f_dst, gm, layer["radial"], drop)

    for n, blk in enumerate(net.down_blocks):
        idx = ops.fps(x, b, net.pool_ratio[n], random_start=not net.deterministic)
        x_dst = ops.gather_rows(x, idx)
        b_dst = b.index_select(0, idx)
        f_dst = project_if_mismatch(blk["pool_proj"], A.GatherFn.apply(f, idx))
        g = ops.radius_csr(x, x_dst, [net.radius[n]], b_src=b, b_dst=b_dst, excl_mode=1, excl=idx, max_num_neighbors=1000)
        f_new = run(blk["pool_layer"], f, f_dst, net._geom(x, x_dst, g))
        graphs.append(("pool", n, idx, x, b))
        f, x, b = f_new, x_dst, b_dst
        outs.append((f, x, b))
        g = ops.radius_csr(x, x, [net.radius[n]], b_src=b, b_dst=b, excl_mode=2, max_num_neighbors=1001)
        geom = net._geom(x, x, g)
        for layer in blk["layer_stack"]:
            f = run(layer, f, f, geom)
            outs.append((f, x, b))
            graphs.append(("self", geom))
        scale_outs.append((f, x, b))
    if net.forward_only:
        # ForwardOnlyFeatureExtractor (forward_only_feature_extractor.py:191-275): no mid / up path, no skips -- every scale
        # outputs its down-path features through project_outputs (mirrors UnetFeatureExtractor.forward's forward_only branch)
        return [FeaturedPoints(x=scale_outs[s][1], f=project_if_mismatch(proj, scale_outs[s][0]), b=scale_outs[s][2], w=None)
                for s, proj in enumerate(net.project_outputs) if s in net.output_scalespace]
    for layer in net.mid_block:
        f = run(layer, f, f, geom)
    f_skip, _, _ = outs.pop()
    s3 = 1.0 / math.sqrt(3)
    f = A.AddScaleFn.apply(f, f_skip, s3)
    ups = []
    for n, blk in enumerate(net.up_blocks):
        for layer in blk["layer_stack"]:
            f_dst, x_dst, b_dst = outs.pop()
            kind = graphs.pop()
            f_dst = A.AddScaleFn.apply(f, f_dst, s3)
            f = run(layer, f, f_dst, kind[1])
            x, b = x_dst, b_dst
        ups.append((f, x, b))
        f_dst, x_dst, b_dst = outs.pop()
        kind = graphs.pop()
        if n != net.n_scales - 1:
            _, scale, idx, x_fine, b_fine = kind
            g = ops.radius_csr(x, x_fine, [net.radius[scale]], b_src=b, b_dst=b_fine, excl_mode=3, excl=idx, max_num_neighbors=1000)
            f = run(blk["unpool_layer"], f, f_dst, net._geom(x, x_fine, g))
            x, b = x_dst, b_dst
    ups = ups[::-1]
    pcds = []
    for s, proj in enumerate(net.project_outputs):
        if s not in net.output_scalespace:
            continue
        fs, xs, bs = ups[s]
        pcds.append(FeaturedPoints(x=xs, f=project_if_mismatch(proj, fs), b=bs, w=None))
    return pcds


# ------------------------------------------------------------------------------------------------ tensor field + head
def tensor_field(field, query_x: torch.Tensor, query_b: torch.Tensor, keys: List[FeaturedPoints],
                 time_emb: Optional[List[torch.Tensor]], rows_per_time: int, pos_grad: bool = False) -> torch.Tensor:
    """MultiscaleTensorField.forward (multiscale_tensor_field.py:192-260) -> (n_query, F).  ``pos_grad``: the edge geometry
    carries a gradient back to ``query_x`` (EbmScoreModelHead.forward); the graph itself (which edges exist) is piecewise
    constant in the coordinates, as in the reference (torch_cluster.radius returns indices)."""
    x_src = torch.cat([p.x for p in keys], dim=0).contiguous()
    f_src = torch.cat([p.f for p in keys], dim=0)
    w_src = torch.cat([p.w for p in keys], dim=0) if field.gnn_block_init.use_src_point_attn else None
    b_src = torch.cat([p.b for p in keys], dim=0).contiguous()
    off = [0]
    for p in keys:
        off.append(off[-1] + p.x.shape[0])
    radii = field.r_cluster_multiscale
    xq = query_x.contiguous()
    g = ops.radius_csr(x_src, xq, radii, src_off=off, b_src=b_src, b_dst=query_b.contiguous(), max_num_neighbors=1000)
    ns = field.r_mincut_nonscalar_sh
    E = g.n_edges
    if pos_grad:
        if E == 0:
            raise NotImplementedError("position gradient with an empty query graph")
        length, sh, logit = A.EdgeGeomFn.apply(query_x, x_src, g, radii, off, (0.2 * ns, 1.0 * ns))
    else:
        length, sh, logit = ops.edge_geom(x_src, xq, g, radii=radii, src_off=off, ns_cut=(0.2 * ns, 1.0 * ns), want_logit=True)
    bounds = g.row_ptr[::g.n_dst][: field.n_scales + 1].tolist()        # edge range of every scale (one host read)
    if field._sin_freq is None or field._sin_freq.device != xq.device:
        half = field.length_emb_dim // 2
        field._sin_freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(1000.0) / (half - 1))).to(xq.device)
    scalars = []
    for s, gp in enumerate(field.graph_parsers):
        e0, e1 = bounds[s], bounds[s + 1]
        if e1 == e0:
            continue
        ln = length[e0:e1].contiguous()
        if gp.r is not None:
            pm = gp.length_enc.param_module
            emb = A.RbfFn.apply(ln, pm.mean, pm.std_logit, pm.weight_logit, 0.0, 1.0 / gp.r, 0)
        elif pos_grad:
            emb = A.SinusoidFn.apply(ln, field._sin_freq, field.length_emb_dim, 1000.0 / float(field.length_enc_max_r))
        else:
            emb = A.sinusoid(ln, field._sin_freq, field.length_emb_dim, 1000.0 / float(field.length_enc_max_r))
        if time_emb is not None:
            rows = torch.div(g.edge_dst[e0:e1], rows_per_time, rounding_mode="floor").to(torch.int32).contiguous()
            emb = torch.cat([emb, A.GatherFn.apply(time_emb[s], rows)], dim=1)
        pre = field.edge_scalars_pre_linears[s][0]
        scalars.append(A.SiluFn.apply(nn_linear(pre, emb)))
    blk = field.gnn_block_init
    msg_src = linear_rs(blk.linear_src, layer_norm(blk.prenorm_src, f_src))
    if not scalars:
        raise NotImplementedError("training step with an empty query graph")
    edge_scalars = torch.cat(scalars, dim=0)
    w = radial_profile(blk.ga.sep_act.dtp_rad, edge_scalars)
    drop = (field.alpha_drop, field.proj_drop) if field.training else (0.0, 0.0)
    emb = graph_attention(blk.ga, msg_src, None, g, sh[:E], w, logit[:E], drop, w_src)
    skip = emb if blk.skip_2.is_identity else project_if_mismatch(blk.skip_2, emb)
    return A.AddScaleFn.apply(ffn(blk.ffn, layer_norm(blk.post_norm, emb), drop[1]), skip, 1.0)


def score_head(head, Ts: torch.Tensor, keys: List[FeaturedPoints], query: FeaturedPoints, time: torch.Tensor):
    """ScoreModelHead.forward (score_head.py:142-211) with gradients."""
    Ts = Ts.contiguous()
    nT, nQ = len(Ts), len(query.x)
    dev = Ts.device
    half = head.time_emb_mlp[0] // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(head.time_enc_n) / (half - 1))).to(dev)
    t_enc = A.sinusoid(time.contiguous(), freq, head.time_emb_mlp[0], head.time_enc_n / head.max_time)
    time_emb = []
    for mlp in head.time_mlps_multiscale:
        h = A.SiluFn.apply(nn_linear(mlp[0], t_enc))
        time_emb.append(nn_linear(mlp[2], h))
    qx, qf = query.x.contiguous(), query.f
    irr = head.irreps_query_edf.m
    xq, _ = ops.query_transform(Ts, qx, qf.detach().contiguous(), irr)
    fq = A.QueryTransformFn.apply(Ts, qx, qf, irr)
    bq = query.b.unsqueeze(0).expand(nT, -1).reshape(-1).contiguous()
    key_f = tensor_field(head.key_tensor_field, xq, bq, keys, time_emb, nQ)
    ys = []
    for tp in (head.lin_vel_tp, head.ang_vel_tp):
        o = A.ScoreTpFn.apply(fq, key_f, tp.dtp.tp.weight, head.irreps_key_edf.m)
        y = linear_rs(tp.lin, o)
        ys.append(A.GateFn.apply(y, tp.lin.irreps_out.m))
    ang, lin = A.AssembleFn.apply(ys[0], ys[1], Ts, qx, query.w, head.n_irreps_prescore, head.lin_mult)
    return ang, lin


# ------------------------------------------------------------------------------------------------ place configs' query model
def keypoint_extractor(mod, input_points: FeaturedPoints) -> FeaturedPoints:
    """KeypointExtractor.forward (keypoint_extractor.py:139-197) with gradients: own UNet -> bbox + FPS query points (no
    gradient: indices) -> feature field and weight field at the query points -> weight head."""
    if mod.weight_mult_logit is not None or not mod.use_sigmoid:
        raise NotImplementedError("training path: weight_activation='sigmoid', weight_mult=None (every shipped place config)")
    keys = unet_forward(mod.feature_extractor, input_points)
    with torch.no_grad():
        q = mod.get_query_points(input_points)
    f = tensor_field(mod.tensor_field, q.x, q.b, keys, None, 1)
    wf = tensor_field(mod.weight_field, q.x, q.b, keys, None, 1)
    ln, lin = mod.weight_post[0], mod.weight_post[2]
    h = A.SiluFn.apply(A.LayerNormFn.apply(wf, ln.weight, ln.bias, (ln.normalized_shape[0], 0, 0), ln.eps))
    w = A.SigmoidFn.apply(nn_linear(lin, h)).reshape(-1)
    return FeaturedPoints(x=q.x, f=f, b=q.b, w=w)


# ------------------------------------------------------------------------------------------------ energy-based head
def ebm_score(head, Ts: torch.Tensor, keys: List[FeaturedPoints], query: FeaturedPoints):
    """EbmScoreModelHead.forward (score_head_ebm.py:192-222), first order (the reference's inference mode): the gradient of
    log P = -energy w.r.t. the transformed query coordinates x' = R x + p and the rotated query features f' = D(R) f comes from
    the adjoint kernels (autograd only orders them); the pull-back to the pose -- what the reference obtains by differentiating
    through quaternion_to_matrix / the Euler-angle Wigner matrices and contracting with L(q) -- is closed form in
    dedf_ebm_pose_grad:  ang_a = sum_q [x_q x R^T g_x + sum_u f_u x R^T g_u + sum_u (X_a f_u) . D2^T g_u],  lin = R^T sum_q g_x."""
    Ts = Ts.detach().to(torch.float32).contiguous()
    nT, nQ = len(Ts), len(query.x)
    irr = head.irreps_query_edf.m
    qx, qf = query.x.detach().contiguous(), query.f.detach().contiguous()
    with torch.enable_grad():
        xq, fq = ops.query_transform(Ts, qx, qf, irr)
        xq, fq = xq.requires_grad_(True), fq.requires_grad_(True)
        bq = query.b.unsqueeze(0).expand(nT, -1).reshape(-1).contiguous()
        key_f = tensor_field(head.key_tensor_field, xq, bq, [FeaturedPoints(x=p.x.detach(), f=p.f.detach(), b=p.b, w=None if p.w is None else p.w.detach()) for p in keys],
                             None, nQ, pos_grad=True)
        energy = A.EbmEnergyFn.apply(key_f, fq, query.w.detach().contiguous(), nT, nQ, head.energy_rescale_factor)
        g_x, g_f = torch.autograd.grad(-energy.sum(), (xq, fq))
    ang = torch.empty(nT, 3, dtype=torch.float32, device=Ts.device)
    lin = torch.empty(nT, 3, dtype=torch.float32, device=Ts.device)
    g_x, g_f = g_x.contiguous(), g_f.contiguous()
    ops._call("dedf_ebm_pose_grad", A.ptr(Ts), nT, nQ, A.L.int_array(irr), A.ptr(qx), A.ptr(qf), A.ptr(g_x), A.ptr(g_f),
              float(head.ang_mult), float(head.lin_mult), A.ptr(ang), A.ptr(lin), A.stream())
    return ang, lin
