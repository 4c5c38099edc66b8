"""MultiscaleTensorField on the CUDA path.

Mirrors /root/reference/diffusion_edf/multiscale_tensor_field.py:16-260 and the edge
encoders of graph_parser.py:17-345 (RadiusBipartite / InfiniteBipartite).  Per call:
one multi-scale radius search (all scales in one launch; edges ordered (scale, dst, src)
exactly like the reference's per-scale concatenation), one geometry kernel, one fused
"length embedding -> pre-linear" launch, one RadialProfile launch and the block.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
from torch import nn

from . import _lib as L
from . import ops
from .block import EquiformerBlock
from .gnn_data import FeaturedPoints
from .irreps import Irreps
from .layers import GaussianRadialBasis, pack_tc, tc_mlp_ok


class _GraphParser(nn.Module):
    """Parameter / buffer holder of RadiusBipartite (finite r) or InfiniteBipartite (r = None)."""

    def __init__(self, r: Optional[float], length_enc_dim: int, length_enc_max_r: Optional[float]):
        super().__init__()
        self.r = None if r is None else float(r)
        self.register_buffer("cutoff_eps", torch.tensor(1e-12))
        if r is not None:
            self.length_enc = GaussianRadialBasis(dim=length_enc_dim, max_val=self.r)
        else:
            assert length_enc_max_r is not None
            self.length_enc = None          # SinusoidalPositionEmbeddings has no parameters
            self.length_enc_max_r = float(length_enc_max_r)


class MultiscaleTensorField(nn.Module):
    def __init__(self, irreps_input, irreps_output, irreps_sh, num_heads: int, fc_neurons: List[int], length_emb_dim: int,
                 irreps_query, r_cluster_multiscale: List[Optional[float]], edge_context_emb_dim: Optional[int],
                 r_mincut_nonscalar_sh: Optional[float] = None, length_enc_max_r: Optional[float] = None,
                 n_scales: Optional[int] = None, n_layers: int = 1, irreps_mlp_mid=3, attn_type: str = "mlp",
                 alpha_drop: float = 0.1, proj_drop: float = 0.1, drop_path_rate: float = 0.0,
                 use_src_point_attn: bool = False, use_dst_point_attn: bool = False, cutoff_method: str = "edge_attn"):
        super().__init__()
        self.irreps_input, self.irreps_output = Irreps(irreps_input), Irreps(irreps_output)
        self.irreps_sh = Irreps(irreps_sh)
        if self.irreps_sh.m not in ((1, 1, 1), (1, 1, 0)):
            raise NotImplementedError("irreps_sh must be the l <= 2 (or l <= 1) spherical harmonics 1x0e+1x1e(+1x2e)")
        if irreps_query is not None:
            raise NotImplementedError("query (dst) features are not used by the shipped configs (query_time_encoding: False)")
        if cutoff_method != "edge_attn" or attn_type != "mlp" or n_layers != 1:
            raise NotImplementedError("only cutoff_method='edge_attn', attn_type='mlp', n_layers=1 are implemented")
        if drop_path_rate and float(drop_path_rate) > 0.0:
            raise NotImplementedError("drop_path_rate > 0 (GraphDropPath, gnn_block.py:156,205-214) is not implemented; every shipped config uses 0.0")
        self.use_dst_feature = False
        self.alpha_drop, self.proj_drop = float(alpha_drop), float(proj_drop)       # train mode only (train_path.py)
        self.num_heads = num_heads
        fc_neurons = list(fc_neurons)
        self.length_emb_dim, self.context_emb_dim = length_emb_dim, edge_context_emb_dim
        ctx = self.context_emb_dim or 0            # None: no context embedding (KeypointExtractor's fields)
        if fc_neurons[0] == -1:
            fc_neurons[0] = self.length_emb_dim + ctx
        assert fc_neurons[0] == self.length_emb_dim + ctx
        self.fc_neurons = fc_neurons
        self.r_cluster_multiscale = list(r_cluster_multiscale)
        self.n_scales = len(self.r_cluster_multiscale)
        assert n_scales is None or n_scales == self.n_scales
        assert self.n_scales <= L.MAX_SCALES
        if r_mincut_nonscalar_sh is None:
            r_mincut_nonscalar_sh = 0.01 * self.r_cluster_multiscale[0]
        self.r_mincut_nonscalar_sh = float(r_mincut_nonscalar_sh)
        seen_inf = False
        self.graph_parsers = nn.ModuleList()
        self.edge_scalars_pre_linears = nn.ModuleList()
        for r in self.r_cluster_multiscale:
            assert not (seen_inf and r is not None), "finite radius after the infinite one"
            seen_inf = seen_inf or r is None
            self.graph_parsers.append(_GraphParser(r, length_emb_dim, length_enc_max_r))
            self.edge_scalars_pre_linears.append(nn.Sequential(nn.Linear(fc_neurons[0], fc_neurons[0]), nn.SiLU()))
        self.length_enc_max_r = length_enc_max_r
        self.n_layers = n_layers
        self.gnn_block_init = EquiformerBlock(irreps_src=self.irreps_input, irreps_dst=self.irreps_input,
                                              irreps_emb=self.irreps_input, irreps_output=self.irreps_output,
                                              irreps_edge_attr=irreps_sh, num_heads=num_heads, fc_neurons=fc_neurons,
                                              irreps_mlp_mid=irreps_mlp_mid, use_dst_feature=False, skip_connection=True,
                                              bias=True, use_src_point_attn=use_src_point_attn,
                                              use_dst_point_attn=use_dst_point_attn, use_edge_weights=True)
        self.gnn_blocks = nn.ModuleList()
        # irreps outside the fused kernels' family (BASELINE config C1: 16x0e+8x1e, l <= 1 harmonics): same parameters, un-fused kernels
        self.fused_family = self.gnn_block_init.ga.fused_family
        self._pre_cache = (None, None)
        self._pre_tc = None
        self._pre_tc16 = None
        self._sin_freq = None

    # ------------------------------------------------------------------ packed params
    def packed_prelinear(self):
        """Transposed halves of the per-scale pre-linears: W_len^T (len_dim, K) and W_time^T (ctx_dim, K), bias."""
        key = tuple((m[0].weight.data_ptr(), m[0].weight._version, m[0].bias._version) for m in self.edge_scalars_pre_linears)
        if self._pre_cache[0] != key:
            with torch.no_grad():
                ld = self.length_emb_dim
                wl = [m[0].weight.detach()[:, :ld].t().contiguous() for m in self.edge_scalars_pre_linears]
                wt = [m[0].weight.detach()[:, ld:].t().contiguous() for m in self.edge_scalars_pre_linears]
                b = [m[0].bias.detach().contiguous() for m in self.edge_scalars_pre_linears]
                # tensor-core layout of the length halves, all scales back to back (dedf_mlp_desc.pre_w_tc)
                dims = [ld] + list(self.gnn_block_init.ga.sep_act.dtp_rad.ch_list)
                self._pre_tc = torch.cat([pack_tc(w) for w in wl]).contiguous() if tc_mlp_ok(dims) else None
                self._pre_tc16 = (torch.cat([pack_tc(w, f16=True) for w in wl]).contiguous()
                                  if self._pre_tc is not None and ld % 16 == 0 else None)
            self._pre_cache = (key, (wl, wt, b))
        return self._pre_cache[1]

    # ------------------------------------------------------------------ scene-side precomputation
    def encode_sources(self, input_points_multiscale: Sequence[FeaturedPoints]):
        """Pose-independent part: concatenated key coordinates / batch ids and linear_src(prenorm_src(f))."""
        assert len(input_points_multiscale) == self.n_scales
        x = torch.cat([p.x for p in input_points_multiscale], dim=0).contiguous()
        f = torch.cat([p.f for p in input_points_multiscale], dim=0).contiguous()
        b = torch.cat([p.b for p in input_points_multiscale], dim=0).contiguous()
        off = [0]
        for p in input_points_multiscale:
            off.append(off[-1] + p.x.shape[0])
        msg = self.gnn_block_init.source_messages(f)
        if self.gnn_block_init.use_src_point_attn:          # PointAttentiveScoreModel: source weights ride along (gnn_data.py:220-234)
            assert all(isinstance(p.w, torch.Tensor) for p in input_points_multiscale), "source-point attention needs FeaturedPoints.w"
            return x, b, off, msg, torch.cat([p.w for p in input_points_multiscale], dim=0).contiguous()
        return x, b, off, msg

    # ------------------------------------------------------------------ forward
    def forward(self, query_points: FeaturedPoints, input_points_multiscale: List[FeaturedPoints],
                context_emb=None, max_neighbors: int = 1000, *, time_rows: Optional[torch.Tensor] = None,
                rows_per_time: int = 1, sources=None, edge_capacity: Optional[int] = None) -> FeaturedPoints:
        """``time_rows`` (n_scales, n_rb, K): time half of the pre-linear from dedf_time_embed; the row used by an
        edge is ``edge_dst // rows_per_time`` (clamped), i.e. nQ consecutive query nodes share a pose's time."""
        if context_emb is not None:
            raise NotImplementedError("pass the time rows produced by ScoreModelHead (time_rows=...), not context_emb")
        if not self.fused_family:
            return self._forward_unfused(query_points, input_points_multiscale, max_neighbors)
        assert (time_rows is not None) == (self.context_emb_dim is not None)
        srcs = sources if sources is not None else self.encode_sources(input_points_multiscale)
        x_src, b_src, src_off, msg_src = srcs[:4]
        xq = query_points.x.contiguous()
        radii = self.r_cluster_multiscale
        g = ops.radius_csr(x_src, xq, radii, src_off=src_off, b_src=b_src, b_dst=query_points.b.contiguous(),
                           max_num_neighbors=max_neighbors, capacity=edge_capacity)
        ns = self.r_mincut_nonscalar_sh
        length, sh, logit = ops.edge_geom(x_src, xq, g, radii=radii, src_off=src_off, ns_cut=(0.2 * ns, 1.0 * ns),
                                          want_logit=True)
        out = self.attend_edges(g, length, sh, logit, srcs, time_rows, rows_per_time)
        return FeaturedPoints(x=query_points.x, f=out, b=query_points.b, w=query_points.w)

    def _forward_unfused(self, query_points: FeaturedPoints, input_points_multiscale: List[FeaturedPoints], max_neighbors: int) -> FeaturedPoints:
        """Any even-parity l <= 2 irreps: the field composed from the un-fused CUDA primitives of the training path (train_path.py:
        radius search, edge geometry, length embedding, pre-linear, RadialProfile, table-driven depthwise tensor products, block-
        diagonal linears, gate, segment softmax-reduce, proj, LN, FFN) -- forward only.  BASELINE config C1 runs here."""
        if self.context_emb_dim is not None:
            raise NotImplementedError("edge context embeddings with irreps outside 2G x0e + G x1e + G/2 x2e")
        if max_neighbors != 1000:
            raise NotImplementedError("max_neighbors other than the reference's default (1000)")
        from . import train_path
        with torch.no_grad():
            f = train_path.tensor_field(self, query_points.x, query_points.b, list(input_points_multiscale), None, 1)
        return FeaturedPoints(x=query_points.x, f=f, b=query_points.b, w=query_points.w)

    def forward_poses(self, Ts: torch.Tensor, query_pcd: FeaturedPoints, sources, *, time_rows: Optional[torch.Tensor] = None,
                      max_neighbors: int = 1000, edge_capacity: Optional[int] = None, step: Optional[torch.Tensor] = None,
                      rows_all: Optional[torch.Tensor] = None, static_sources: bool = False) -> torch.Tensor:
        """The field at the pose-transformed query points ``T_t . x_q`` for every pose t and query point q, rows ordered (t, q):
        ``forward(TransformPcd(query_pcd, Ts).flatten(), ...)`` with the point transform, the radius search, the CSR and the edge
        geometry fused into one launch (dedf_head_front).  -> (n_t * n_q, F) features.  Denoise loop: ``step`` (device step index)
        and ``rows_all`` (n_scales, n_steps, K) select this step's time rows, which are written to ``time_rows`` (n_scales, 1, K)."""
        x_src, b_src, src_off, msg_src = sources[:4]
        ns = self.r_mincut_nonscalar_sh
        g, length, sh, logit, _ = ops.head_front(Ts, query_pcd.x.contiguous(), query_pcd.b.contiguous(), x_src, b_src, src_off,
                                                 self.r_cluster_multiscale, (0.2 * ns, 1.0 * ns), max_num_neighbors=max_neighbors,
                                                 capacity=edge_capacity, step=step, rows_all=rows_all,
                                                 rows_cur=time_rows if rows_all is not None else None, static_sources=static_sources)
        return self.attend_edges(g, length, sh, logit, sources, time_rows, query_pcd.x.shape[0])

    def attend_edges(self, g: ops.Csr, length: torch.Tensor, sh: torch.Tensor, logit: torch.Tensor, srcs, time_rows: Optional[torch.Tensor],
                     rows_per_time: int) -> torch.Tensor:
        """Edge scalars -> per-edge tensor-product weights -> the attention block, on a finished graph."""
        assert (time_rows is not None) == (self.context_emb_dim is not None)
        x_src, b_src, src_off, msg_src = srcs[:4]
        w_src = srcs[4] if len(srcs) > 4 else None
        # length embedding + pre-linear (+ time rows) -> h0 ; RadialProfile -> per-edge TP weights
        wl, _, pre_bias = self.packed_prelinear()
        K = self.fc_neurons[0]
        E = max(1, g.n_edges)
        dev = length.device
        d = L.MlpDesc()
        d.mode = L.MLP_IN_FIELD
        d.n_edges_dev = L.ptr(g.n_edges_dev, torch.int32)
        d.length = L.ptr(length)
        d.n_scales, d.n_dst = self.n_scales, g.n_dst
        d.row_ptr, d.edge_dst = L.ptr(g.row_ptr, torch.int32), L.ptr(g.edge_dst, torch.int32)
        keep = []
        for s, gp in enumerate(self.graph_parsers):
            if gp.r is not None:
                pm = gp.length_enc.param_module
                t = (pm.mean.detach().reshape(-1), pm.std_logit.detach().reshape(-1), pm.weight_logit.detach().reshape(-1))
                keep.append(t)
                d.enc_mean[s], d.enc_std_logit[s], d.enc_weight_logit[s] = L.ptr(t[0]), L.ptr(t[1]), L.ptr(t[2])
                d.enc_r[s] = gp.r
            else:
                d.enc_r[s] = -1.0
            d.pre_w[s] = L.ptr(wl[s])
        if time_rows is None:
            # without a context embedding the per-scale bias rides in row_bias (one row per scale)
            rb = torch.stack(pre_bias, dim=0).unsqueeze(1).contiguous()      # (n_scales, 1, K)
            keep.append(rb)
            d.row_bias = L.ptr(rb)
            d.n_rb, d.rb_div = 1, 1
        d.enc_max_r = float(self.length_enc_max_r) if self.length_enc_max_r is not None else 1.0
        d.enc_n = 1000.0
        half = self.length_emb_dim // 2
        if self._sin_freq is None or self._sin_freq.device != dev:
            # SinusoidalPositionEmbeddings frequency table (radial_func.py:310-312), fp32 exp on the host
            self._sin_freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(1000.0) / (half - 1))).to(dev)
        d.enc_freq = L.ptr(self._sin_freq)
        if time_rows is not None:
            d.row_bias = L.ptr(time_rows)          # W_time t_emb + b, per pose
            d.n_rb, d.rb_div = time_rows.shape[1], rows_per_time
        d.dims[0], d.dims[1] = self.length_emb_dim, K
        d.flags[0] = 2
        rad = self.gnn_block_init.ga.sep_act.dtp_rad
        w = torch.empty(E, self.gnn_block_init.ga.sep_act.numel, dtype=torch.float32, device=dev)
        if ops.USE_TC_MLP and self._pre_tc is not None:
            # tensor cores: pre-linear + RadialProfile in ONE launch (the (E, K) pre-linear output stays on chip)
            f16 = self._pre_tc16 is not None and rad.f16_ok()
            rad.fill_desc(d, 1, f16=f16)
            d.pre_w_tc = L.ptr(self._pre_tc16 if f16 else self._pre_tc)
            d.out = L.ptr(w)
            ops.edge_mlp_tc(d, g.n_edges)
        else:
            h0 = torch.empty(E, K, dtype=torch.float32, device=dev)
            d.n_layers = 1
            d.out = L.ptr(h0)
            ops.edge_mlp(d, g.n_edges)
            d2 = L.MlpDesc()
            d2.mode = L.MLP_IN_ROWS
            d2.n_edges_dev = L.ptr(g.n_edges_dev, torch.int32)
            d2.x_in = L.ptr(h0)
            rad.fill_desc(d2, 0)
            d2.out = L.ptr(w)
            ops.edge_mlp(d2, g.n_edges)

        return self.gnn_block_init(msg_src, g, sh, w, logit, w_src)
