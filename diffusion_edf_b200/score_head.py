"""ScoreModelHead on the CUDA path.  Mirrors /root/reference/diffusion_edf/score_head.py:18-246
(constructor kwargs, ``forward`` / ``warmup`` signatures, ``jittable`` attribute, parameter names)."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import _lib as L
from . import ops
from .gnn_data import FeaturedPoints, TransformPcd
from .irreps import Irreps
from .layers import LinearRS, _DTP
from .multiscale_tensor_field import MultiscaleTensorField


class _ScoreTP(nn.Module):
    """SeparableFCTP(irreps, irreps -> 1x0e + NV x1e, internal weights, gate): parameter holder
    (equiformer/graph_attention_transformer.py:60-135; weight layout SURVEY.md App. E.2)."""

    def __init__(self, irr: Irreps, n_vec: int):
        super().__init__()
        m0, m1, m2 = irr.m
        shapes = [(m0, m0), (m0, m1), (m1, m0), (m1, m1), (m1, m1), (m1, m2), (m2, m1), (m2, m2), (m2, m2)]
        scales = torch.cat([torch.full((a * b,), 1.0 / (b ** 0.5)) for a, b in shapes])   # uvu: fan_in = mul2
        self.dtp = _DTP(sum(a * b for a, b in shapes), True, scales)
        d0, d1 = m0 + m1 + m2, m0 + 3 * m1 + 2 * m2
        self.lin = LinearRS(Irreps((d0, d1, 0)), Irreps((1 + n_vec, n_vec, 0)))
        self._shapes = shapes

    def packed_dtp(self) -> torch.Tensor:
        """dtp.tp.weight with every path block transposed to [mul2][mul1] (the layout dedf_score_tp reads coalesced)."""
        w, off, out = self.dtp.tp.weight.detach(), 0, []
        for a, b in self._shapes:
            out.append(w[off:off + a * b].view(a, b).t().contiguous().view(-1))
            off += a * b
        return torch.cat(out).contiguous()


class ScoreModelHead(nn.Module):
    jittable: bool = False      # the CUDA path is not TorchScript-able (and does not need to be)

    def __init__(self, max_time: float, time_emb_mlp: List[int], key_tensor_field_kwargs: Dict, irreps_query_edf,
                 lin_mult: float, ang_mult: float, time_enc_n: float = 10000.0, edge_time_encoding: bool = False,
                 query_time_encoding: bool = True):
        super().__init__()
        if not edge_time_encoding or query_time_encoding:
            raise NotImplementedError("only edge_time_encoding=True / query_time_encoding=False (all shipped configs)")
        self.lin_mult, self.ang_mult = float(lin_mult), float(ang_mult)
        self.max_time, self.time_enc_n = float(max_time), float(time_enc_n)
        self.edge_time_encoding, self.query_time_encoding = edge_time_encoding, query_time_encoding
        self.n_scales = key_tensor_field_kwargs.get("n_scales", len(key_tensor_field_kwargs["r_cluster_multiscale"]))
        self.time_emb_mlp = list(time_emb_mlp)
        if len(self.time_emb_mlp) != 3:
            raise NotImplementedError("time_emb_mlp must have 3 entries [enc, hidden, out]")
        self.time_mlps_multiscale = nn.ModuleList()
        for _ in range(self.n_scales):
            self.time_mlps_multiscale.append(nn.Sequential(nn.Linear(time_emb_mlp[0], time_emb_mlp[1]), nn.SiLU(inplace=True),
                                                           nn.Linear(time_emb_mlp[1], time_emb_mlp[2])))
        self.query_time_mlp = None
        self.time_emb_dim = time_emb_mlp[-1]
        # the reference mutates the kwargs dict in place (score_head.py:81-93); keep that contract
        assert "irreps_query" not in key_tensor_field_kwargs and "edge_context_emb_dim" not in key_tensor_field_kwargs
        key_tensor_field_kwargs["irreps_query"] = None
        key_tensor_field_kwargs["edge_context_emb_dim"] = self.time_emb_mlp[-1]
        self.key_tensor_field = MultiscaleTensorField(**key_tensor_field_kwargs)
        self.irreps_key_edf = self.key_tensor_field.irreps_output
        self.key_edf_dim = self.irreps_key_edf.dim
        self.irreps_query_edf = Irreps(irreps_query_edf)
        self.query_edf_dim = self.irreps_query_edf.dim
        if self.irreps_query_edf != self.irreps_key_edf:
            raise NotImplementedError("query and key EDF irreps must agree")
        self.query_transform = TransformPcd(irreps=self.irreps_query_edf)
        self.n_irreps_prescore = (self.irreps_query_edf.m[1] + self.irreps_key_edf.m[1]) // 2
        self.lin_vel_tp = _ScoreTP(self.irreps_key_edf, self.n_irreps_prescore)
        self.ang_vel_tp = _ScoreTP(self.irreps_key_edf, self.n_irreps_prescore)
        self._time_cache = (None, None)
        self._tp_cache = (None, None)

    # ------------------------------------------------------------------ packed params
    def _time_desc(self) -> L.TimeDesc:
        mods = list(self.time_mlps_multiscale) + [m[0] for m in self.key_tensor_field.edge_scalars_pre_linears]
        key = tuple((p.data_ptr(), p._version) for m in mods for p in m.parameters())
        if self._time_cache[0] != key:
            with torch.no_grad():
                _, wt, bp = self.key_tensor_field.packed_prelinear()
                d = L.TimeDesc()
                d.max_time, d.enc_n = self.max_time, self.time_enc_n
                half = self.time_emb_mlp[0] // 2
                dev = self.time_mlps_multiscale[0][0].weight.device
                # frequency table tabulated exactly like radial_func.py:310-312 (fp32 exp on the host)
                freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(self.time_enc_n) / (half - 1))).to(dev)
                d.enc_freq = L.ptr(freq)
                d.enc_dim, d.h_dim, d.e_dim = self.time_emb_mlp
                d.out_dim, d.n_scales = self.key_tensor_field.fc_neurons[0], self.n_scales
                keep = [freq]
                for s, mlp in enumerate(self.time_mlps_multiscale):
                    w1, b1 = mlp[0].weight.detach().t().contiguous(), mlp[0].bias.detach().contiguous()
                    w2, b2 = mlp[2].weight.detach().t().contiguous(), mlp[2].bias.detach().contiguous()
                    keep += [w1, b1, w2, b2, wt[s], bp[s]]
                    d.W1[s], d.b1[s], d.W2[s], d.b2[s] = L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2)
                    d.Wp[s], d.bp[s] = L.ptr(wt[s]), L.ptr(bp[s])
            self._time_cache = (key, (d, keep))
        return self._time_cache[1][0]

    def _tp_packed(self):
        key = tuple((p.data_ptr(), p._version) for m in (self.lin_vel_tp, self.ang_vel_tp) for p in m.parameters())
        if self._tp_cache[0] != key:
            with torch.no_grad():
                Wd, Wl0, Wl1, bl = [], [], [], []
                for tp in (self.lin_vel_tp, self.ang_vel_tp):
                    (w0, w1, _), b = tp.lin.packed()
                    Wd.append(tp.packed_dtp()); Wl0.append(w0); Wl1.append(w1); bl.append(b)
            self._tp_cache = (key, (Wd, Wl0, Wl1, bl))
        return self._tp_cache[1]

    # ------------------------------------------------------------------ forward
    def forward(self, Ts: torch.Tensor, key_pcd_multiscale: List[FeaturedPoints], query_pcd: FeaturedPoints,
                time: torch.Tensor, *, sources=None, shared_time: bool = False,
                edge_capacity: Optional[int] = None, time_rows: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        assert Ts.ndim == 2 and Ts.shape[-1] == 7, f"{Ts.shape}"
        assert time.ndim == 1 and (len(time) == len(Ts) or shared_time), f"{time.shape}"
        assert query_pcd.f.ndim == 2 and query_pcd.f.shape[-1] == self.query_edf_dim, f"{query_pcd.f.shape}"
        assert isinstance(query_pcd.w, torch.Tensor)
        Ts = Ts.contiguous()
        nT, nQ = len(Ts), len(query_pcd.x)
        # time encoding -> per-scale MLP -> time half of the edge pre-linear: (n_scales, nT or 1, K)
        if time_rows is None:        # (sample() precomputes the rows of the whole schedule and passes the current step's)
            time_rows = ops.time_embed(self._time_desc(), (time[:1] if shared_time else time).contiguous())
        qx, qf = query_pcd.x.contiguous(), query_pcd.f.contiguous()
        Wd, Wl0, Wl1, bl = self._tp_packed()
        field = self.key_tensor_field
        if ops.USE_HEAD_FRONT and self._front_ok(key_pcd_multiscale, sources):
            # fused front (points transform + radius search + CSR + geometry) and D(q) psi inside the score kernel
            srcs = sources if sources is not None else field.encode_sources(key_pcd_multiscale)
            key_f = field.forward_poses(Ts, query_pcd, srcs, time_rows=time_rows, edge_capacity=edge_capacity)
            return ops.score_tp_step(Ts, qf, key_f, qx, query_pcd.w.contiguous(), self.irreps_key_edf.m, Wd, Wl0, Wl1, bl,
                                     self.n_irreps_prescore, self.lin_mult)
        # query transform: x' = R x + t, f' = D(q) f
        xq, fq = ops.query_transform(Ts, qx, qf, self.irreps_query_edf.m)
        bq = query_pcd.b.unsqueeze(0).expand(nT, -1).reshape(-1).contiguous()
        flat = FeaturedPoints(x=xq, f=fq, b=bq, w=None)
        out = field(query_points=flat, input_points_multiscale=key_pcd_multiscale,
                    time_rows=time_rows, rows_per_time=nQ, sources=sources, edge_capacity=edge_capacity)
        ang, lin = ops.score_tp(Ts, fq, out.f, qx, query_pcd.w.contiguous(), self.irreps_key_edf.m, Wd, Wl0, Wl1, bl,
                                self.n_irreps_prescore, self.lin_mult)
        return ang, lin

    def _front_ok(self, key_pcd_multiscale, sources) -> bool:
        """dedf_head_front stages the concatenated scene scales in shared memory (<= 200 KB: ~12 k points)."""
        n = sources[0].shape[0] if sources is not None else sum(p.x.shape[0] for p in key_pcd_multiscale)
        return n * 16 + 16 <= 200 * 1024

    def denoise_step(self, T32: torch.Tensor, query_pcd: FeaturedPoints, sources, edge_capacity: int, overflow: torch.Tensor, state) -> None:
        """One step of ScoreModelBase.sample on device-resident state (denoise.StepState): six launches, no host values.
        score_model_base.py:174-199 with score_head.py:142-211 inside."""
        field = self.key_tensor_field
        Wd, Wl0, Wl1, bl = self._tp_packed()
        with ops.use_plan(None):
            ns = field.r_mincut_nonscalar_sh
            g, length, sh, logit, _ = ops.head_front(T32, query_pcd.x, query_pcd.b, sources[0], sources[1], sources[2],
                                                     field.r_cluster_multiscale, (0.2 * ns, 1.0 * ns), capacity=edge_capacity,
                                                     overflow=overflow, step=state.counter, rows_all=state.rows_all,
                                                     rows_cur=state.rows_cur, static_sources=True)
            key_f = field.attend_edges(g, length, sh, logit, sources, state.rows_cur, query_pcd.x.shape[0])
            ops.score_tp_step(T32, query_pcd.f, key_f, query_pcd.x, query_pcd.w, self.irreps_key_edf.m, Wd, Wl0, Wl1, bl,
                              self.n_irreps_prescore, self.lin_mult, state=state)

    def time_rows_for(self, times: torch.Tensor) -> torch.Tensor:
        """Time half of the edge pre-linear for a whole schedule in one launch: (n_scales, len(times), K)."""
        return ops.time_embed(self._time_desc(), times.to(torch.float32).contiguous())

    def warmup(self, Ts, key_pcd_multiscale, query_pcd, time):
        return self.forward(Ts=Ts, key_pcd_multiscale=key_pcd_multiscale, query_pcd=query_pcd, time=time)
