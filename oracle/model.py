"""CPU restatement of the Diffusion-EDF score network.
Oracle = test infrastructure only; never imported by the product package.

Module / attribute names mirror the reference so ``state_dict`` keys agree
(SURVEY.md App. C).  Follows (relative to /root/reference/diffusion_edf):
  graph_attention.py:16-122 (GraphAttentionMLP), :138-273 (GraphAttentionMLP2)
  gnn_block.py:21-57 (FeedForwardNetwork), :65-218 (EquiformerBlock, head)
  block.py:64-174 (EquiformerBlock, UNet; incl. the discarded-norm quirk :149-153)
  graph_parser.py:17-224 (edge encoding), :229-286, :291-345
  multiscale_tensor_field.py:16-260, wigner.py:17-19,119-125,232-283
  gnn_data.py:80-113, score_head.py:18-211, keypoint_extractor.py:22-47
  connectivity.py:8-76, utils.py:26-47, unet_feature_extractor.py:19-417
  score_model_base.py:22-225, multiscale_score_model.py:25-117
"""
from __future__ import annotations

import copy

import math
from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple, Union

import torch
from torch import nn

from . import encoders as enc
from . import graph as G
from . import so3
from .irreps import Irreps, multiply_irreps, sort_even_first
from .nn import (EquivariantLayerNormV2, FullyConnectedTensorProductRescale,
                 FullyConnectedTensorProductRescaleSwishGate, LinearRS, ProjectIfMismatch, SeparableFCTP,
                 act_consts, heads2vec, smooth_leaky_relu, vec2heads)


class FeaturedPoints(NamedTuple):
    x: torch.Tensor
    f: torch.Tensor
    b: torch.Tensor
    w: Optional[torch.Tensor] = None


class GraphEdge(NamedTuple):
    edge_src: torch.Tensor
    edge_dst: torch.Tensor
    edge_length: Optional[torch.Tensor] = None
    edge_attr: Optional[torch.Tensor] = None
    edge_scalars: Optional[torch.Tensor] = None
    edge_weights: Optional[torch.Tensor] = None
    edge_logits: Optional[torch.Tensor] = None


def get_mul_0(irreps: Irreps) -> int:
    return irreps.count(0, 1)


# ==========================================================================
# attention
# ==========================================================================
class _GraphAttentionCore(nn.Module):
    """Shared arithmetic of GraphAttentionMLP / GraphAttentionMLP2."""

    def _build(self, irreps_in, irreps_mid, irreps_edge_attr, irreps_out, irreps_head, fc_neurons, num_heads):
        self.irreps_head, self.num_heads = Irreps(irreps_head), num_heads
        heads, _, _ = sort_even_first(self.irreps_head * num_heads)
        heads = heads.simplify()
        self.irreps_attn_heads = heads
        mul_alpha = get_mul_0(heads)
        self.mul_alpha_head = mul_alpha // num_heads
        assert self.mul_alpha_head * num_heads == mul_alpha
        self.sep_act = SeparableFCTP(irreps_in, irreps_edge_attr, irreps_mid, fc_neurons,
                                     use_activation=True, internal_weights=False)
        self.sep_alpha = LinearRS(self.sep_act.dtp.irreps_out, Irreps(f"{mul_alpha}x0e"))
        self.sep_value = SeparableFCTP(irreps_mid, irreps_edge_attr, heads, None,
                                       use_activation=False, internal_weights=True)
        self.alpha_dot = nn.Parameter(torch.randn(1, num_heads, self.mul_alpha_head))
        nn.init.xavier_uniform_(self.alpha_dot)
        self.proj = LinearRS(heads, irreps_out)
        self.c_slrelu = act_consts()["slrelu"]

    def _attend(self, message, edge_dst, edge_attr, edge_scalars, n_nodes_dst, pre_logit, post_attn):
        weight = self.sep_act.dtp_rad(edge_scalars)
        message = self.sep_act.dtp(message, edge_attr, weight)
        log_alpha = self.sep_alpha(message).reshape(len(message), self.num_heads, self.mul_alpha_head)
        value = self.sep_act.gate(self.sep_act.lin(message))
        value = self.sep_value(value, edge_attr=edge_attr, edge_scalars=edge_scalars)
        log_alpha = self.c_slrelu * smooth_leaky_relu(log_alpha)
        log_alpha = torch.einsum("ehk,hk->eh", log_alpha, self.alpha_dot.squeeze(0))
        if pre_logit is not None:
            log_alpha = log_alpha + pre_logit.unsqueeze(-1)
        value = vec2heads(value, self.irreps_head, self.num_heads)
        log_Z = G.scatter_logsumexp(log_alpha, edge_dst, n_nodes_dst)
        alpha = torch.exp(log_alpha - log_Z[edge_dst])
        if post_attn is not None:
            alpha = alpha * post_attn.unsqueeze(-1)
        attn = G.scatter_sum(value * alpha.unsqueeze(-1), edge_dst, n_nodes_dst)
        return self.proj(heads2vec(attn, self.irreps_head))


class GraphAttentionMLP(_GraphAttentionCore):
    def __init__(self, irreps_emb, irreps_edge_attr, irreps_node_output, fc_neurons, irreps_head, num_heads,
                 alpha_drop=0.1, proj_drop=0.1):
        super().__init__()
        self._build(Irreps(irreps_emb), Irreps(irreps_emb), Irreps(irreps_edge_attr), Irreps(irreps_node_output),
                    irreps_head, fc_neurons, num_heads)

    def forward(self, message, edge_dst, edge_attr, edge_scalars, n_nodes_dst):
        return self._attend(message, edge_dst, edge_attr, edge_scalars, n_nodes_dst, None, None)


class GraphAttentionMLP2(_GraphAttentionCore):
    def __init__(self, irreps_input, irreps_edge_attr, irreps_output, fc_neurons, num_heads,
                 alpha_drop=0.1, proj_drop=0.1):
        super().__init__()
        irreps_input = Irreps(irreps_input)
        head = multiply_irreps(irreps_input, 1 / num_heads)
        self._build(irreps_input, irreps_input, Irreps(irreps_edge_attr), Irreps(irreps_output), head,
                    fc_neurons, num_heads)

    def forward(self, message, graph_edge: GraphEdge, n_nodes_dst, edge_pre_attn_logit=None, edge_post_attn=None):
        return self._attend(message, graph_edge.edge_dst, graph_edge.edge_attr, graph_edge.edge_scalars,
                            n_nodes_dst, edge_pre_attn_logit, edge_post_attn)


class FeedForwardNetwork(nn.Module):
    def __init__(self, irreps_node_input, irreps_node_output, irreps_mlp_mid):
        super().__init__()
        one = Irreps("1x0e")
        self.fctp_1 = FullyConnectedTensorProductRescaleSwishGate(Irreps(irreps_node_input), one, Irreps(irreps_mlp_mid))
        self.fctp_2 = FullyConnectedTensorProductRescale(Irreps(irreps_mlp_mid), one, Irreps(irreps_node_output))

    def forward(self, x):
        ones = torch.ones_like(x[:, 0:1])
        return self.fctp_2(self.fctp_1(x, ones), ones)


def _mlp_mid(irreps_emb: Irreps, mult: Union[int, Irreps]) -> Irreps:
    if isinstance(mult, int):
        return sort_even_first(irreps_emb * mult)[0].simplify()
    return Irreps(mult)


class UnetEquiformerBlock(nn.Module):
    """block.py:64-174.  NOTE the quirk at :149-153: the layer norms are computed
    and discarded -- the linears see the un-normalised inputs."""

    def __init__(self, irreps_src, irreps_dst, irreps_edge_attr, irreps_head, num_heads, fc_neurons,
                 irreps_mlp_mid=3, src_bias=False, dst_bias=True, **_unused):
        super().__init__()
        self.irreps_src, self.irreps_dst = Irreps(irreps_src), Irreps(irreps_dst)
        self.irreps_emb = self.irreps_dst
        self.norm_1_src = EquivariantLayerNormV2(self.irreps_src)
        self.linear_src = LinearRS(self.irreps_src, self.irreps_emb, bias=src_bias)
        self.norm_1_dst = EquivariantLayerNormV2(self.irreps_dst)
        self.linear_dst = LinearRS(self.irreps_dst, self.irreps_emb, bias=dst_bias)
        self.ga = GraphAttentionMLP(self.irreps_emb, irreps_edge_attr, self.irreps_dst, fc_neurons, irreps_head, num_heads)
        self.norm_2 = EquivariantLayerNormV2(self.irreps_dst)
        self.ffn = FeedForwardNetwork(self.irreps_dst, self.irreps_dst, _mlp_mid(self.irreps_emb, irreps_mlp_mid))

    def forward(self, node_input_src, node_input_dst, batch_dst, edge_src, edge_dst, edge_attr, edge_scalars):
        message = self.linear_src(node_input_src)[edge_src] + self.linear_dst(node_input_dst)[edge_dst]
        feat = self.ga(message, edge_dst, edge_attr, edge_scalars, len(node_input_dst))
        out = node_input_dst + feat
        return out + self.ffn(self.norm_2(out))


class EquiformerBlock(nn.Module):
    """gnn_block.py:65-218 (the score-head block)."""

    def __init__(self, irreps_src, irreps_dst, irreps_edge_attr, num_heads, fc_neurons, irreps_emb=None,
                 irreps_output=None, irreps_mlp_mid=3, use_dst_feature=True, skip_connection=True, bias=True,
                 use_src_point_attn=False, use_dst_point_attn=False, use_edge_weights=True, **_unused):
        super().__init__()
        self.irreps_src, self.irreps_dst = Irreps(irreps_src), Irreps(irreps_dst)
        self.irreps_emb = Irreps(irreps_emb) if irreps_emb is not None else self.irreps_dst
        self.irreps_output = Irreps(irreps_output) if irreps_output is not None else self.irreps_dst
        self.use_dst_feature, self.use_edge_weights = use_dst_feature, use_edge_weights
        self.use_src_point_attn = use_src_point_attn
        assert not use_dst_point_attn
        self.skip_1 = self.skip_2 = None
        if skip_connection:
            if use_dst_feature:
                self.skip_1 = ProjectIfMismatch(self.irreps_dst, self.irreps_emb, bias=True, layernorm=False)
            self.skip_2 = ProjectIfMismatch(self.irreps_emb, self.irreps_output, bias=True, layernorm=False)
        self.prenorm_src = EquivariantLayerNormV2(self.irreps_src)
        if use_dst_feature:
            self.linear_src = LinearRS(self.irreps_src, self.irreps_emb, bias=False)
            self.prenorm_dst = EquivariantLayerNormV2(self.irreps_dst)
            self.linear_dst = LinearRS(self.irreps_dst, self.irreps_emb, bias=True)
        else:
            self.linear_src = LinearRS(self.irreps_src, self.irreps_emb, bias=True)
            self.prenorm_dst = self.linear_dst = None
        self.ga = GraphAttentionMLP2(self.irreps_emb, irreps_edge_attr, self.irreps_emb, fc_neurons, num_heads)
        self.post_norm = EquivariantLayerNormV2(self.irreps_emb, affine=bias)
        self.ffn = FeedForwardNetwork(self.irreps_emb, self.irreps_output, _mlp_mid(self.irreps_emb, irreps_mlp_mid))

    def forward(self, src_points: FeaturedPoints, dst_points: FeaturedPoints, graph_edge: GraphEdge) -> FeaturedPoints:
        message = self.linear_src(self.prenorm_src(src_points.f))[graph_edge.edge_src]
        if self.prenorm_dst is not None:
            message = message + self.linear_dst(self.prenorm_dst(dst_points.f))[graph_edge.edge_dst]
        pre = graph_edge.edge_logits if self.use_edge_weights else None
        post = src_points.w[graph_edge.edge_src] if self.use_src_point_attn else None
        emb = self.ga(message, graph_edge, len(dst_points.x), pre, post)
        if self.skip_1 is not None:
            emb = emb + self.skip_1(dst_points.f)
        out = self.ffn(self.post_norm(emb))
        if self.skip_2 is not None:
            out = out + self.skip_2(emb)
        return FeaturedPoints(x=dst_points.x, f=out, b=dst_points.b, w=dst_points.w)


# ==========================================================================
# edge encoding / graph parsers
# ==========================================================================
def cutoff_irreps(f, cutoff_nonscalar, irreps: Irreps):
    """irreps_utils.py:20-63 with only ``cutoff_nonscalar`` set."""
    if cutoff_nonscalar is None:
        return f
    out, off = [], 0
    for m, l, _ in irreps:
        d = m * (2 * l + 1)
        blk = f[..., off:off + d]
        out.append(blk * cutoff_nonscalar[..., None] if l != 0 else blk)
        off += d
    return torch.cat(out, dim=-1)


class GraphEdgeEncoderBase(nn.Module):
    def __init__(self, r_cutoff, irreps_sh, length_enc, r_mincut_nonscalar_sh, cutoff_eps=1e-12, sh_cutoff=False):
        super().__init__()
        assert not sh_cutoff, "cutoff_method='sh' is not used by any shipped config"
        self.register_buffer("cutoff_eps", torch.tensor(cutoff_eps))
        self.length_enc = length_enc
        self.edge_cutoff_ranges = None if r_cutoff is None else (None, None, 0.8 * float(r_cutoff), 1.0 * float(r_cutoff))
        self.nonscalar_sh_cutoff_ranges = None
        if r_mincut_nonscalar_sh is not None:
            self.nonscalar_sh_cutoff_ranges = (0.2 * float(r_mincut_nonscalar_sh), 1.0 * float(r_mincut_nonscalar_sh), None, None)
        self.irreps_sh = Irreps(irreps_sh)

    def _encode_edges(self, x_src, x_dst, edge_src, edge_dst, fill_edge_weights: Optional[float] = None) -> GraphEdge:
        vec = x_src.index_select(0, edge_src) - x_dst.index_select(0, edge_dst)
        length = vec.norm(dim=1, p=2)
        edge_cutoff = None if self.edge_cutoff_ranges is None else enc.soft_square_cutoff_2(length, self.edge_cutoff_ranges)
        cut_ns = None if self.nonscalar_sh_cutoff_ranges is None else enc.soft_square_cutoff_2(length, self.nonscalar_sh_cutoff_ranges)
        scalars = self.length_enc(length) if self.length_enc is not None else None
        sh = so3.spherical_harmonics(self.irreps_sh.lmax, vec, normalize=True)
        sh = cutoff_irreps(sh, cut_ns, self.irreps_sh)
        if edge_cutoff is None:
            if fill_edge_weights is None:
                w = logit = None
            else:
                w = torch.ones_like(length) * fill_edge_weights
                logit = torch.ones_like(length) * math.log(fill_edge_weights)
        else:
            w = torch.max(edge_cutoff, self.cutoff_eps)
            logit = torch.log(w)
        return GraphEdge(edge_src, edge_dst, length, sh, scalars, w, logit)


class InfiniteBipartite(GraphEdgeEncoderBase):
    def __init__(self, irreps_sh, r_mincut_nonscalar_sh, length_enc_dim, length_enc_max_r, sh_cutoff=False,
                 fill_edge_weights=False):
        le = enc.SinusoidalPositionEmbeddings(dim=length_enc_dim, max_val=float(length_enc_max_r), n=1000.0)
        super().__init__(None, irreps_sh, le, r_mincut_nonscalar_sh, sh_cutoff=sh_cutoff)
        self.fill_edge_weights = 1.0 if fill_edge_weights else None

    def forward(self, src: FeaturedPoints, dst: FeaturedPoints, max_neighbors=None) -> GraphEdge:
        es, ed = torch.meshgrid(torch.arange(len(src.x)), torch.arange(len(dst.x)), indexing="ij")
        return self._encode_edges(src.x, dst.x, es.reshape(-1), ed.reshape(-1), self.fill_edge_weights)


class RadiusBipartite(GraphEdgeEncoderBase):
    def __init__(self, r_cutoff, irreps_sh, length_enc_dim, r_mincut_nonscalar_sh, sh_cutoff=False):
        self.r_cluster = float(r_cutoff)
        le = enc.GaussianRadialBasis(dim=length_enc_dim, max_val=self.r_cluster)
        super().__init__(r_cutoff, irreps_sh, le, r_mincut_nonscalar_sh, sh_cutoff=sh_cutoff)

    def forward(self, src: FeaturedPoints, dst: FeaturedPoints, max_neighbors: int = 1000) -> GraphEdge:
        e = G.radius(src.x, dst.x, self.r_cluster, src.b, dst.b, max_neighbors)
        return self._encode_edges(src.x, dst.x, e[1], e[0])


def _cat_opt(a, b):
    if a is None or b is None:
        assert a is None and b is None
        return None
    return torch.cat([a, b], dim=0)


class MultiscaleTensorField(nn.Module):
    def __init__(self, irreps_input, irreps_output, irreps_sh, num_heads, fc_neurons, length_emb_dim, irreps_query,
                 r_cluster_multiscale, edge_context_emb_dim, r_mincut_nonscalar_sh=None, length_enc_max_r=None,
                 n_scales=None, n_layers=1, irreps_mlp_mid=3, attn_type="mlp", alpha_drop=0.1, proj_drop=0.1,
                 drop_path_rate=0.0, use_src_point_attn=False, use_dst_point_attn=False, cutoff_method="edge_attn"):
        super().__init__()
        self.irreps_input, self.irreps_output, self.irreps_sh = Irreps(irreps_input), Irreps(irreps_output), Irreps(irreps_sh)
        self.use_dst_feature = irreps_query is not None
        self.irreps_query = Irreps(irreps_query) if irreps_query is not None else None
        fc_neurons = list(fc_neurons)
        self.length_emb_dim, self.context_emb_dim = length_emb_dim, edge_context_emb_dim
        if fc_neurons[0] == -1:
            fc_neurons[0] = length_emb_dim + (edge_context_emb_dim or 0)
        assert fc_neurons[0] == length_emb_dim + (edge_context_emb_dim or 0)
        assert cutoff_method == "edge_attn"
        self.r_cluster_multiscale = list(r_cluster_multiscale)
        self.n_scales = len(self.r_cluster_multiscale)
        if r_mincut_nonscalar_sh is None:
            r_mincut_nonscalar_sh = 0.01 * self.r_cluster_multiscale[0]
        self.graph_parsers, self.edge_scalars_pre_linears = nn.ModuleList(), nn.ModuleList()
        fill = False
        for r in self.r_cluster_multiscale:
            if r is None:
                self.graph_parsers.append(InfiniteBipartite(self.irreps_sh, r_mincut_nonscalar_sh, length_emb_dim,
                                                            length_enc_max_r, fill_edge_weights=fill))
            else:
                self.graph_parsers.append(RadiusBipartite(r, self.irreps_sh, length_emb_dim, r_mincut_nonscalar_sh))
                fill = True
            self.edge_scalars_pre_linears.append(nn.Sequential(nn.Linear(fc_neurons[0], fc_neurons[0]), nn.SiLU()))
        self.n_layers = n_layers
        assert n_layers == 1, "only n_layers=1 is used by the shipped configs"
        self.gnn_block_init = EquiformerBlock(
            irreps_src=self.irreps_input, irreps_dst=self.irreps_query if self.use_dst_feature else self.irreps_input,
            irreps_emb=self.irreps_input, irreps_output=self.irreps_output, irreps_edge_attr=self.irreps_sh,
            num_heads=num_heads, fc_neurons=fc_neurons, irreps_mlp_mid=irreps_mlp_mid,
            use_dst_feature=self.use_dst_feature, skip_connection=True, bias=True,
            use_src_point_attn=use_src_point_attn, use_dst_point_attn=use_dst_point_attn, use_edge_weights=True)
        self.gnn_blocks = nn.ModuleList()

    def forward(self, query_points: FeaturedPoints, input_points_multiscale: List[FeaturedPoints],
                context_emb: Optional[List[torch.Tensor]] = None, max_neighbors: int = 1000) -> FeaturedPoints:
        n_total, edges, pts = 0, None, None
        for n, (parser, pre) in enumerate(zip(self.graph_parsers, self.edge_scalars_pre_linears)):
            ip = input_points_multiscale[n]
            ge = parser(src=ip, dst=query_points, max_neighbors=max_neighbors)
            s = ge.edge_scalars
            if self.context_emb_dim is not None:
                s = torch.cat([s, context_emb[n].index_select(0, ge.edge_dst)], dim=-1)
            s = pre(s)
            ge = ge._replace(edge_scalars=s, edge_src=ge.edge_src + n_total)
            n_total += len(ip.x)
            if edges is None:
                edges, pts = ge, ip
            else:
                edges = GraphEdge(*[_cat_opt(a, b) for a, b in zip(edges, ge)])
                pts = FeaturedPoints(torch.cat([pts.x, ip.x]), torch.cat([pts.f, ip.f]), torch.cat([pts.b, ip.b]), None)
        return self.gnn_block_init(src_points=pts, dst_points=query_points, graph_edge=edges)


# ==========================================================================
# query transform + score head
# ==========================================================================
def transform_features(irreps: Irreps, feature: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """wigner.py:257-283: (nQ, D) x (nT, 4) -> (nT, nQ, D) through YXY Euler angles and J."""
    if irreps.lmax == 0:
        return feature.expand(len(q), -1, -1)
    q = enc.standardize_quaternion(q / torch.norm(q, dim=-1, keepdim=True))
    ang = enc.matrix_to_euler_yxy(enc.quaternion_to_matrix(q)).T
    a, b, c = ang[0], ang[1], ang[2]
    out = []
    for (m, l, _), sl in zip(irreps, irreps.slices()):
        f = feature[..., sl]
        if l == 0:
            out.append(f.expand(len(a), len(f), f.shape[-1]))
        else:
            D = so3.wigner_D_euler(l, a, b, c)
            t = torch.einsum("tij,qmj->tqmi", D, f.reshape(f.shape[0], -1, 2 * l + 1))
            out.append(t.reshape(t.shape[0], t.shape[1], -1))
    return torch.cat(out, dim=-1)


class SliceAndTransform(nn.Module):
    def __init__(self, l):
        super().__init__()
        self.register_buffer("J", so3.J_matrix(l).to(torch.float32).clone())


class TransformFeatureQuaternion(nn.Module):
    def __init__(self, irreps):
        super().__init__()
        self.irreps = Irreps(irreps)
        self.transforms = nn.ModuleList([SliceAndTransform(l) for _, l, _ in self.irreps])

    def forward(self, feature, q):
        return transform_features(self.irreps, feature, q)


class TransformPcd(nn.Module):
    def __init__(self, irreps):
        super().__init__()
        self.transform_features = TransformFeatureQuaternion(irreps)

    def forward(self, pcd: FeaturedPoints, Ts: torch.Tensor) -> FeaturedPoints:
        f = self.transform_features(pcd.f, Ts[..., :4])
        x = enc.transform_points(pcd.x, Ts)
        w = pcd.w.expand(len(Ts), -1) if pcd.w is not None else None
        return FeaturedPoints(x=x, f=f, b=pcd.b.expand(len(Ts), -1), w=w)


class ScoreModelHead(nn.Module):
    def __init__(self, max_time, time_emb_mlp, key_tensor_field_kwargs, irreps_query_edf, lin_mult, ang_mult,
                 time_enc_n=10000.0, edge_time_encoding=False, query_time_encoding=True):
        super().__init__()
        assert edge_time_encoding and not query_time_encoding, "only the edge-time-encoding variant is used by shipped configs"
        self.lin_mult, self.ang_mult = lin_mult, ang_mult
        kw = dict(key_tensor_field_kwargs)
        self.n_scales = len(kw["r_cluster_multiscale"])
        self.time_emb_mlp = list(time_emb_mlp)
        self.time_enc = enc.SinusoidalPositionEmbeddings(dim=time_emb_mlp[0], max_val=max_time, n=time_enc_n)
        self.time_mlps_multiscale = nn.ModuleList()
        for _ in range(self.n_scales):
            layers = []
            for i in range(1, len(time_emb_mlp)):
                layers.append(nn.Linear(time_emb_mlp[i - 1], time_emb_mlp[i]))
                if i != len(time_emb_mlp) - 1:
                    layers.append(nn.SiLU())
            self.time_mlps_multiscale.append(nn.Sequential(*layers))
        self.time_emb_dim = time_emb_mlp[-1]
        kw["irreps_query"] = None
        kw["edge_context_emb_dim"] = time_emb_mlp[-1]
        self.key_tensor_field = MultiscaleTensorField(**kw)
        self.irreps_key_edf = self.key_tensor_field.irreps_output
        self.irreps_query_edf = Irreps(irreps_query_edf)
        self.query_transform = TransformPcd(self.irreps_query_edf)
        self.n_irreps_prescore = (self.irreps_query_edf.count(1, 1) + self.irreps_key_edf.count(1, 1)) // 2
        pres = Irreps("1x0e") + Irreps(f"{self.n_irreps_prescore}x1e")
        self.lin_vel_tp = SeparableFCTP(self.irreps_key_edf, self.irreps_query_edf, pres, None,
                                        use_activation=True, internal_weights=True)
        self.ang_vel_tp = SeparableFCTP(self.irreps_key_edf, self.irreps_query_edf, pres, None,
                                        use_activation=True, internal_weights=True)

    def forward(self, Ts, key_pcd_multiscale: List[FeaturedPoints], query_pcd: FeaturedPoints, time):
        nT, nQ = len(Ts), len(query_pcd.x)
        time_enc = self.time_enc(time)
        time_embs = [mlp(time_enc).unsqueeze(-2).expand(-1, nQ, -1).reshape(nT * nQ, self.time_emb_dim)
                     for mlp in self.time_mlps_multiscale]
        qt = self.query_transform(pcd=query_pcd, Ts=Ts)
        qf = qt.f.clone().reshape(-1, qt.f.shape[-1])
        flat = FeaturedPoints(x=qt.x.reshape(-1, 3), f=torch.empty_like(qf), b=qt.b.reshape(-1), w=None)
        field = self.key_tensor_field(query_points=flat, input_points_multiscale=key_pcd_multiscale, context_emb=time_embs)
        lin = self.lin_vel_tp(qf, field.f, edge_scalars=None)[..., 1:]
        ang = self.ang_vel_tp(qf, field.f, edge_scalars=None)[..., 1:]
        lin = lin.view(nT, nQ, self.n_irreps_prescore, 3).mean(dim=-2)
        ang = ang.view(nT, nQ, self.n_irreps_prescore, 3).mean(dim=-2)
        qinv = enc.quaternion_invert(Ts[..., :4].unsqueeze(-2))
        lin = enc.quaternion_apply(qinv, lin)
        ang = enc.quaternion_apply(qinv, ang)
        orbital = torch.cross(query_pcd.x.unsqueeze(0) / self.lin_mult, lin, dim=-1)
        w = query_pcd.w
        lin_vel = torch.einsum("q,tqi->ti", w, lin)
        ang_vel = torch.einsum("q,tqi->ti", w, orbital) + torch.einsum("q,tqi->ti", w, ang)
        return ang_vel, lin_vel


class EbmScoreModelHead(nn.Module):
    """Energy-based head (critic), /root/reference/diffusion_edf/score_head_ebm.py:32-222.  ``compute_energy`` is what
    agent.py:163-174 calls to re-rank samples; ``forward`` (the score as the gradient of the energy w.r.t. the pose,
    :192-222) is restated with torch autograd."""

    def __init__(self, max_time, time_emb_mlp, key_tensor_field_kwargs, irreps_query_edf, lin_mult, ang_mult,
                 time_enc_n=10000.0, edge_time_encoding=False, query_time_encoding=True):
        super().__init__()
        assert not edge_time_encoding and not query_time_encoding, "the shipped *_ebm configs use no time encoding"
        self.lin_mult, self.ang_mult = lin_mult, ang_mult
        kw = dict(key_tensor_field_kwargs)
        self.n_scales = len(kw["r_cluster_multiscale"])
        self.time_emb_mlp = list(time_emb_mlp)
        self.time_enc = enc.SinusoidalPositionEmbeddings(dim=time_emb_mlp[0], max_val=max_time, n=time_enc_n)
        self.time_mlps_multiscale = nn.ModuleList()
        for _ in range(self.n_scales):                      # present in the state_dict, unused without time encoding
            layers = []
            for i in range(1, len(time_emb_mlp)):
                layers.append(nn.Linear(time_emb_mlp[i - 1], time_emb_mlp[i]))
                if i != len(time_emb_mlp) - 1:
                    layers.append(nn.SiLU())
            self.time_mlps_multiscale.append(nn.Sequential(*layers))
        kw["irreps_query"] = None
        kw["edge_context_emb_dim"] = None
        self.key_tensor_field = MultiscaleTensorField(**kw)
        self.irreps_key_edf = self.key_tensor_field.irreps_output
        self.irreps_query_edf = Irreps(irreps_query_edf)
        self.query_transform = TransformPcd(self.irreps_query_edf)
        self.register_buffer("q_indices", torch.tensor([[1, 2, 3], [0, 3, 2], [3, 0, 1], [2, 1, 0]], dtype=torch.long), persistent=False)
        self.register_buffer("q_factor", torch.tensor([[-0.5, -0.5, -0.5], [0.5, -0.5, 0.5], [0.5, 0.5, -0.5], [-0.5, 0.5, 0.5]]), persistent=False)
        self.energy_rescale_factor = 1.0 / float(self.irreps_key_edf.dim)

    def compute_energy(self, Ts, key_pcd_multiscale: List[FeaturedPoints], query_pcd: FeaturedPoints, time):
        nT, nQ = len(Ts), len(query_pcd.x)
        qt = self.query_transform(pcd=query_pcd, Ts=Ts)
        qf = qt.f.clone().reshape(-1, qt.f.shape[-1])
        flat = FeaturedPoints(x=qt.x.reshape(-1, 3), f=torch.empty_like(qf), b=qt.b.reshape(-1), w=None)
        field = self.key_tensor_field(query_points=flat, input_points_multiscale=key_pcd_multiscale, context_emb=None)
        energy = (field.f - qf).square().sum(dim=-1) * self.energy_rescale_factor
        return torch.einsum("q,tq->t", query_pcd.w, energy.view(nT, nQ))

    def forward(self, Ts, key_pcd_multiscale: List[FeaturedPoints], query_pcd: FeaturedPoints, time):
        """score_head_ebm.py:192-222: the score is the gradient of log P = -energy w.r.t. the pose, pulled back to the body
        frame: ang = L(q)^T d/dq (right-trivialised quaternion derivative), lin = R(q)^-1 d/dp.  First order only
        (the reference's inference mode, ``create_graph=False``)."""
        with torch.enable_grad():
            T = Ts.detach().clone().requires_grad_(True)
            logp = -self.compute_energy(T, key_pcd_multiscale, query_pcd, time)
            grad = torch.autograd.grad(logp.sum(), T)[0]
        L = T.detach()[..., self.q_indices] * self.q_factor
        ang_vel = torch.einsum("...ia,...i", L, grad[..., :4]) * self.ang_mult
        lin_vel = enc.quaternion_apply(enc.quaternion_invert(T[..., :4].detach()), grad[..., 4:]) * self.lin_mult
        return ang_vel.detach(), lin_vel.detach()


class StaticKeypointModel(nn.Module):
    def __init__(self, keypoint_coords, irreps_output):
        super().__init__()
        kc = torch.tensor(keypoint_coords)
        self.irreps_output = Irreps(irreps_output)
        self.register_buffer("keypoint_coords", kc)
        self.keypoint_features = nn.Parameter(torch.randn(len(kc), self.irreps_output.dim))
        self.keypoint_weights = nn.Parameter(torch.randn(len(kc)))

    def forward(self, input_points: FeaturedPoints) -> FeaturedPoints:
        bu = torch.unique(input_points.b)
        n = len(bu)
        return FeaturedPoints(x=self.keypoint_coords.repeat(n, 1), f=self.keypoint_features.repeat(n, 1),
                              b=bu.repeat(len(self.keypoint_coords)), w=torch.sigmoid(self.keypoint_weights).repeat(n))


class KeypointExtractor(nn.Module):
    """keypoint_extractor.py:50-197 (UnetFeatureExtractor, sigmoid / none weight activation)."""

    def __init__(self, feature_extractor_kwargs, tensor_field_kwargs, keypoint_kwargs, feature_extractor_name="UnetFeatureExtractor",
                 weight_activation="sigmoid", weight_mult=None, deterministic=False):
        super().__init__()
        assert feature_extractor_name == "UnetFeatureExtractor"
        self.pool_ratio = float(keypoint_kwargs["pool_ratio"])
        self.keypoint_bbox = keypoint_kwargs.get("bbox", None)
        self.weight_pre_emb_dim = int(keypoint_kwargs["weight_pre_emb_dim"])
        self.weight_mult_logit = None if weight_mult is None else nn.Parameter(torch.log(torch.exp(torch.tensor(float(weight_mult))) - 1))
        self.feature_extractor = UnetFeatureExtractor(**feature_extractor_kwargs, deterministic=deterministic)
        kw = dict(tensor_field_kwargs)
        kw.update(irreps_input=feature_extractor_kwargs["irreps_output"], irreps_query=None, edge_context_emb_dim=None)
        self.tensor_field = MultiscaleTensorField(**kw)
        kw["irreps_output"] = f"{self.weight_pre_emb_dim}x0e"
        self.weight_field = MultiscaleTensorField(**kw)
        self.weight_post = nn.Sequential(nn.LayerNorm(self.weight_pre_emb_dim), nn.SiLU(), nn.Linear(self.weight_pre_emb_dim, 1),
                                         nn.Sigmoid() if weight_activation == "sigmoid" else nn.Identity())
        self.irreps_output = Irreps(self.tensor_field.irreps_output)

    def get_query_points(self, src: FeaturedPoints) -> FeaturedPoints:
        x, b = src.x, src.b
        if self.keypoint_bbox is not None:
            bb = torch.tensor(self.keypoint_bbox, dtype=x.dtype)
            idx = ((x >= bb[:, 0]) * (x <= bb[:, 1])).all(dim=-1).nonzero().squeeze(-1)
            x, b = x.index_select(0, idx), b.index_select(0, idx)
        sel = G.fps(x, b, self.pool_ratio, random_start=False)
        x, b = x.index_select(0, sel), b.index_select(0, sel)
        return FeaturedPoints(x=x, f=torch.empty_like(x), b=b, w=None)

    def forward(self, input_points: FeaturedPoints, max_neighbors: int = 1000) -> FeaturedPoints:
        keys = self.feature_extractor(input_points)
        q = self.get_query_points(input_points)
        out = self.tensor_field(query_points=q, input_points_multiscale=keys, context_emb=None, max_neighbors=max_neighbors)
        w = self.weight_field(query_points=q, input_points_multiscale=keys, context_emb=None, max_neighbors=max_neighbors).f
        w = self.weight_post(w).squeeze(-1)
        if self.weight_mult_logit is not None:
            w = w * torch.nn.functional.softplus(self.weight_mult_logit)
        return FeaturedPoints(x=out.x, f=out.f, b=out.b, w=w)


# ==========================================================================
# key encoder (UNet)
# ==========================================================================
class _Graph(NamedTuple):
    src: torch.Tensor
    dst: torch.Tensor
    length: torch.Tensor
    attr: torch.Tensor


class ParityInversionSh(nn.Module):
    def __init__(self, irreps):
        super().__init__()
        self.register_buffer("sign", torch.cat([(1.0 if l % 2 == 0 else -1.0) * torch.ones((2 * l + 1) * m)
                                                for m, l, _ in Irreps(irreps)]))

    def forward(self, x):
        return x * self.sign


class _Layer(nn.ModuleDict):
    pass


class UnetFeatureExtractor(nn.Module):
    def __init__(self, irreps_input, irreps_output, irreps_emb, irreps_edge_attr, num_heads, fc_neurons, n_layers,
                 pool_ratio, radius, deterministic=False, pool_method="fps", irreps_mlp_mid=3, attn_type="mlp",
                 alpha_drop=0.1, proj_drop=0.1, drop_path_rate=0.0, n_layers_midstream=2, n_scales=None,
                 output_scalespace=None):
        super().__init__()
        self.irreps_output = Irreps(irreps_output)
        self.irreps_emb = [Irreps(i) for i in irreps_emb]
        self.irreps_edge_attr = [Irreps(i) for i in irreps_edge_attr]
        self.n_scales = len(self.irreps_emb)
        self.num_heads, self.fc_neurons, self.pool_ratio, self.n_layers = num_heads, fc_neurons, pool_ratio, n_layers
        self.deterministic = deterministic
        self.irreps_input = Irreps(irreps_input)
        self.input_emb = LinearRS(self.irreps_input, self.irreps_emb[0], bias=True)
        self.output_scalespace = list(range(self.n_scales)) if output_scalespace is None else \
            [self.n_scales + n if n < 0 else n for n in output_scalespace]
        self.radius = [radius[0]]
        for n, r in enumerate(radius[1:]):           # unet_feature_extractor.py:79-86 (pool_ratio[n-1] quirk)
            self.radius.append(self.radius[-1] / math.sqrt(self.pool_ratio[n - 1]) if r is None else r)
        mid = irreps_mlp_mid if isinstance(irreps_mlp_mid, list) else [irreps_mlp_mid] * self.n_scales
        head = [multiply_irreps(self.irreps_emb[n], 1 / num_heads[n]) for n in range(self.n_scales)]

        def layer(n, src, dst, head_irreps):
            return _Layer({
                "radial": enc.GaussianRadialBasisLayerFiniteCutoff(num_basis=fc_neurons[n][0], cutoff=0.99 * self.radius[n]),
                "gnn": UnetEquiformerBlock(src, dst, self.irreps_edge_attr[n], head_irreps, num_heads[n], fc_neurons[n],
                                           irreps_mlp_mid=mid[n], src_bias=False, dst_bias=True)})

        self.down_blocks = nn.ModuleList()
        for n in range(self.n_scales):
            blk = nn.ModuleDict()
            blk["pool_proj"] = ProjectIfMismatch(self.irreps_emb[max(n - 1, 0)], self.irreps_emb[n])
            blk["pool_layer"] = layer(n, self.irreps_emb[max(n - 1, 0)], self.irreps_emb[n], head[n])
            blk["layer_stack"] = nn.ModuleList([layer(n, self.irreps_emb[n], self.irreps_emb[n], head[n])
                                                for _ in range(n_layers[n] - 1)])
            self.down_blocks.append(blk)
        self.mid_block = nn.ModuleList([layer(self.n_scales - 1, self.irreps_emb[-1], self.irreps_emb[-1], head[-1])
                                        for _ in range(n_layers_midstream)])
        self.up_blocks = nn.ModuleList()
        for n in range(self.n_scales - 1, -1, -1):
            blk = nn.ModuleDict()
            blk["parity_inversion"] = ParityInversionSh(self.irreps_edge_attr[n])
            blk["layer_stack"] = nn.ModuleList([layer(n, self.irreps_emb[n], self.irreps_emb[n], head[n])
                                                for _ in range(n_layers[n] - 1)])
            blk["unpool_layer"] = layer(n, self.irreps_emb[n], self.irreps_emb[max(n - 1, 0)], head[max(n - 1, 0)])
            self.up_blocks.append(blk)
        self.project_outputs = nn.ModuleList([ProjectIfMismatch(self.irreps_emb[n], self.irreps_output)
                                              for n in range(self.n_scales)])

    # ---- connectivity.py ---------------------------------------------------
    def _fps_pool(self, n, x, f, b):
        idx = G.fps(x, b, self.pool_ratio[n], random_start=False)   # oracle == deterministic=True
        x_dst, b_dst = x[idx], b[idx]
        e = G.radius(x, x_dst, self.radius[n], b, b_dst, 1000)
        e_dst, e_src = e[0], e[1]
        keep = idx[e_dst] != e_src
        return f[idx], x_dst, e_src[keep], e_dst[keep], b_dst

    def _radius_graph(self, n, x, b):
        e = G.radius_graph(x, self.radius[n], b, loop=False, max_num_neighbors=1000)
        return e[1], e[0]

    def _geom(self, n, x_src, x_dst, e_src, e_dst) -> _Graph:
        vec = x_src.index_select(0, e_src) - x_dst.index_select(0, e_dst)
        return _Graph(e_src, e_dst, vec.norm(dim=1, p=2), so3.spherical_harmonics(self.irreps_edge_attr[n].lmax, vec))

    @staticmethod
    def _run(layer, f_src, f_dst, b_dst, g: _Graph):
        return layer["gnn"](f_src, f_dst, b_dst, g.src, g.dst, g.attr, layer["radial"](g.length))

    def forward(self, pcd: FeaturedPoints) -> List[FeaturedPoints]:
        x, f, b = pcd.x, self.input_emb(pcd.f), pcd.b
        outs, graphs = [(f, x, b)], []
        for n, blk in enumerate(self.down_blocks):
            f_dst, x_dst, e_src, e_dst, b_dst = self._fps_pool(n, x, f, b)
            f_dst = blk["pool_proj"](f_dst)
            g = self._geom(n, x, x_dst, e_src, e_dst)
            f = self._run(blk["pool_layer"], f, f_dst, b_dst, g)
            x, b = x_dst, b_dst
            outs.append((f, x, b)); graphs.append(g)
            e_src, e_dst = self._radius_graph(n, x, b)
            g = self._geom(n, x, x, e_src, e_dst)
            for layer in blk["layer_stack"]:
                f = self._run(layer, f, f, b, g)
                outs.append((f, x, b)); graphs.append(g)
        for layer in self.mid_block:
            f = self._run(layer, f, f, b, g)
        f_skip, _, _ = outs.pop()
        f = (f + f_skip) / math.sqrt(3)
        ups = []
        for n, blk in enumerate(self.up_blocks):
            for layer in blk["layer_stack"]:
                f_dst, x_dst, b_dst = outs.pop()
                g = graphs.pop()
                g = _Graph(g.dst, g.src, g.length, blk["parity_inversion"](g.attr))
                f_dst = (f + f_dst) / math.sqrt(3)
                f = self._run(layer, f, f_dst, b_dst, g)
                x, b = x_dst, b_dst
            ups.append((f, x, b))
            f_dst, x_dst, b_dst = outs.pop()
            g = graphs.pop()
            g = _Graph(g.dst, g.src, g.length, blk["parity_inversion"](g.attr))
            if n != self.n_scales - 1:
                f = self._run(blk["unpool_layer"], f, f_dst, b_dst, g)
                x, b = x_dst, b_dst
        ups = ups[::-1]
        return [FeaturedPoints(x=ups[s][1], f=proj(ups[s][0]), b=ups[s][2], w=None)
                for s, proj in enumerate(self.project_outputs) if s in self.output_scalespace]


class ForwardOnlyFeatureExtractor(UnetFeatureExtractor):
    """/root/reference/diffusion_edf/forward_only_feature_extractor.py:19-275: the UNet's down path only (no mid-stream, no up
    path); scale n outputs the features after its layer stack.  Same parameter names as the down path of the UNet."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        del self.mid_block, self.up_blocks

    def forward(self, pcd: FeaturedPoints) -> List[FeaturedPoints]:
        x, f, b = pcd.x, self.input_emb(pcd.f), pcd.b
        outs = []
        for n, blk in enumerate(self.down_blocks):
            f_dst, x_dst, e_src, e_dst, b_dst = self._fps_pool(n, x, f, b)
            f_dst = blk["pool_proj"](f_dst)
            f = self._run(blk["pool_layer"], f, f_dst, b_dst, self._geom(n, x, x_dst, e_src, e_dst))
            x, b = x_dst, b_dst
            e_src, e_dst = self._radius_graph(n, x, b)
            g = self._geom(n, x, x, e_src, e_dst)
            for layer in blk["layer_stack"]:
                f = self._run(layer, f, f, b, g)
            outs.append((f, x, b))
        return [FeaturedPoints(x=outs[s][1], f=proj(outs[s][0]), b=outs[s][2], w=None)
                for s, proj in enumerate(self.project_outputs) if s in self.output_scalespace]


# ==========================================================================
# top-level model
# ==========================================================================
class MultiscaleScoreModel(nn.Module):
    def __init__(self, query_model: str, score_head_kwargs: Dict, key_kwargs: Dict, query_kwargs: Dict,
                 deterministic: bool = False):
        super().__init__()
        self.register_buffer("q_indices", torch.tensor([[1, 2, 3], [0, 3, 2], [3, 0, 1], [2, 1, 0]], dtype=torch.long), persistent=False)
        self.register_buffer("q_factor", torch.tensor([[-0.5, -0.5, -0.5], [0.5, -0.5, 0.5], [0.5, 0.5, -0.5], [-0.5, 0.5, 0.5]]), persistent=False)
        fe_cls = {"UnetFeatureExtractor": UnetFeatureExtractor, "ForwardOnlyFeatureExtractor": ForwardOnlyFeatureExtractor}[key_kwargs["feature_extractor_name"]]
        self.key_model = fe_cls(**key_kwargs["feature_extractor_kwargs"], deterministic=deterministic)
        if query_model == "StaticKeypointModel":
            self.query_model = StaticKeypointModel(**query_kwargs)
        else:
            assert query_model == "KeypointExtractor"
            self.query_model = KeypointExtractor(**query_kwargs, deterministic=deterministic)
        kw = dict(score_head_kwargs["key_tensor_field_kwargs"])
        kw.update(irreps_input=self.key_model.irreps_output, use_src_point_attn=False, use_dst_point_attn=False)
        head_cls = EbmScoreModelHead if score_head_kwargs.get("ebm", False) else ScoreModelHead
        self.score_head = head_cls(max_time=float(score_head_kwargs["max_time"]),
                                         time_emb_mlp=score_head_kwargs["time_emb_mlp"],
                                         key_tensor_field_kwargs=kw, irreps_query_edf=self.query_model.irreps_output,
                                         lin_mult=float(score_head_kwargs["lin_mult"]), ang_mult=float(score_head_kwargs["ang_mult"]),
                                         edge_time_encoding=score_head_kwargs["edge_time_encoding"],
                                         query_time_encoding=score_head_kwargs["query_time_encoding"])
        self.lin_mult, self.ang_mult = self.score_head.lin_mult, self.score_head.ang_mult

    def get_key_pcd_multiscale(self, pcd):
        return self.key_model(pcd)

    def get_query_pcd(self, pcd):
        return self.query_model(pcd)

    def forward(self, Ts, time, key_pcd, query_pcd, debug=False):
        key_ms = self.get_key_pcd_multiscale(key_pcd)
        q = self.get_query_pcd(query_pcd)
        score = self.score_head(Ts=Ts, key_pcd_multiscale=key_ms, query_pcd=q, time=time)
        return score, ((key_ms, q) if debug else None)

    def get_train_loss(self, Ts, time, key_pcd, query_pcd, target_ang_score, target_lin_score):
        key_ms = self.get_key_pcd_multiscale(key_pcd)
        q = self.get_query_pcd(query_pcd)
        ang, lin = self.score_head(Ts=Ts, key_pcd_multiscale=key_ms, query_pcd=q, time=time)
        t_ang = target_ang_score * torch.sqrt(time[..., None]) * self.ang_mult
        t_lin = target_lin_score * torch.sqrt(time[..., None]) * self.lin_mult
        ang_loss = torch.sum(torch.square(t_ang - ang), dim=-1).mean(dim=-1)
        lin_loss = torch.sum(torch.square(t_lin - lin), dim=-1).mean(dim=-1)
        return ang_loss + lin_loss, {"ang_score": ang.detach(), "lin_score": lin.detach()}

    @torch.no_grad()
    def sample(self, T_seed, scene_pcd_multiscale, grasp_pcd, diffusion_schedules, N_steps, timesteps,
               temperatures=1.0, log_t_schedule=True, time_exponent_temp=0.5, time_exponent_alpha=0.5,
               noise: Optional[torch.Tensor] = None):
        """score_model_base.py:110-204.  ``noise`` (sum(N_steps), nT, 6) float64 standard
        normals (ang then lin) may be injected so that two implementations can be
        compared with temperature > 0; ``None`` draws them with torch.randn."""
        if isinstance(temperatures, (int, float)):
            temperatures = [float(temperatures)] * len(diffusion_schedules)
        dtype = T_seed.dtype
        T = T_seed.clone().detach().double()
        temps = torch.tensor(temperatures, dtype=torch.float64)
        scheds = torch.tensor(diffusion_schedules, dtype=torch.float64)
        Ts, step = [T.clone()], 0
        for n, sch in enumerate(scheds):
            if log_t_schedule:
                ts = torch.logspace(torch.log(sch[0]), torch.log(sch[1]), N_steps[n], base=torch.e, dtype=torch.float64).unsqueeze(-1)
            else:
                ts = torch.linspace(sch[0], sch[1], N_steps[n], dtype=torch.float64).unsqueeze(-1)
            for i in range(len(ts)):
                t = ts[i]
                temperature = temps[n] * torch.pow(t, time_exponent_temp)
                a_ang = (self.ang_mult ** 2) * torch.pow(t, time_exponent_alpha) * timesteps[n]
                a_lin = (self.lin_mult ** 2) * torch.pow(t, time_exponent_alpha) * timesteps[n]
                ang, lin = self.score_head(Ts=T.view(-1, 7).type(dtype), key_pcd_multiscale=scene_pcd_multiscale,
                                           query_pcd=grasp_pcd, time=t.repeat(len(T)).type(dtype))
                ang = ang.double() / (self.ang_mult * torch.sqrt(t))
                lin = lin.double() / (self.lin_mult * torch.sqrt(t))
                if noise is None:
                    z_ang, z_lin = torch.randn_like(ang), torch.randn_like(lin)
                else:
                    z_ang, z_lin = noise[step, :, :3], noise[step, :, 3:]
                ang_disp = (a_ang / 2) * ang + torch.sqrt(temperature * a_ang) * z_ang
                lin_disp = (a_lin / 2) * lin + torch.sqrt(temperature * a_lin) * z_lin
                L = T[..., self.q_indices] * self.q_factor.double()
                q, x = T[..., :4], T[..., 4:]
                dq = torch.einsum("...ij,...j->...i", L, ang_disp)
                dx = enc.quaternion_apply(q, lin_disp)
                q = enc.normalize_quaternion(q + dq)
                T = torch.cat([q, x + dx], dim=-1)
                step += 1
                Ts.append(T.clone())
        Ts.append(T.clone())
        return torch.stack(Ts, dim=0)


class PointAttentiveScoreModel(MultiscaleScoreModel):
    """/root/reference/diffusion_edf/point_attentive_score_model.py:20-99: KeypointExtractor on the key side, one key "scale",
    source-point attention in the score head.  forward / get_train_loss / sample are inherited."""

    def __init__(self, query_model: str, score_head_kwargs: Dict, key_kwargs: Dict, query_kwargs: Dict, deterministic: bool = False):
        nn.Module.__init__(self)
        self.register_buffer("q_indices", torch.tensor([[1, 2, 3], [0, 3, 2], [3, 0, 1], [2, 1, 0]], dtype=torch.long), persistent=False)
        self.register_buffer("q_factor", torch.tensor([[-0.5, -0.5, -0.5], [0.5, -0.5, 0.5], [0.5, 0.5, -0.5], [-0.5, 0.5, 0.5]]), persistent=False)
        self.key_model = KeypointExtractor(**copy.deepcopy(key_kwargs), deterministic=deterministic)
        if query_model == "StaticKeypointModel":
            self.query_model = StaticKeypointModel(**query_kwargs)
        else:
            assert query_model == "KeypointExtractor"
            self.query_model = KeypointExtractor(**copy.deepcopy(query_kwargs), deterministic=deterministic)
        kw = dict(score_head_kwargs["key_tensor_field_kwargs"])
        kw.update(irreps_input=self.key_model.irreps_output, use_src_point_attn=True, use_dst_point_attn=False)
        self.score_head = ScoreModelHead(max_time=float(score_head_kwargs["max_time"]), time_emb_mlp=score_head_kwargs["time_emb_mlp"],
                                         key_tensor_field_kwargs=kw, irreps_query_edf=self.query_model.irreps_output,
                                         lin_mult=float(score_head_kwargs["lin_mult"]), ang_mult=float(score_head_kwargs["ang_mult"]),
                                         edge_time_encoding=score_head_kwargs["edge_time_encoding"],
                                         query_time_encoding=score_head_kwargs["query_time_encoding"])
        self.lin_mult, self.ang_mult = self.score_head.lin_mult, self.score_head.ang_mult

    def get_key_pcd_multiscale(self, pcd):
        return [self.key_model(pcd)]
