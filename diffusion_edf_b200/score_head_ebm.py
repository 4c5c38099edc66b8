"""EbmScoreModelHead (critic) on the CUDA path.  Mirrors /root/reference/diffusion_edf/score_head_ebm.py:32-222: same
constructor kwargs and parameter names.  ``compute_energy`` -- what agent.py:163-174 calls under no_grad to re-rank the
sampled poses -- runs on the kernels of the tensor field plus one energy kernel.  ``forward`` (the score as the gradient
of -energy w.r.t. the pose, :192-222) runs the differentiable (un-fused) field kernels of train_path.py with the adjoints
w.r.t. the query coordinates (csrc/train.cu "position gradients") in inference mode; in train mode (create_graph=True: a
double backward) it raises."""
from __future__ import annotations

from typing import Dict, List

import torch
from torch import nn

from . import ops
from .gnn_data import FeaturedPoints, TransformPcd
from .irreps import Irreps
from .multiscale_tensor_field import MultiscaleTensorField


class EbmScoreModelHead(nn.Module):
    jittable: bool = False

    def __init__(self, max_time: float, time_emb_mlp: List[int], key_tensor_field_kwargs: Dict, irreps_query_edf,
                 lin_mult: float, ang_mult: float, time_enc_n: float = 10000.0, edge_time_encoding: bool = False,
                 query_time_encoding: bool = True):
        super().__init__()
        if edge_time_encoding or query_time_encoding:
            raise NotImplementedError("EbmScoreModelHead: only edge_time_encoding=False / query_time_encoding=False (all shipped *_ebm configs)")
        self.lin_mult, self.ang_mult = float(lin_mult), float(ang_mult)
        self.max_time, self.time_enc_n = float(max_time), float(time_enc_n)
        self.edge_time_encoding, self.query_time_encoding = False, False
        self.n_scales = key_tensor_field_kwargs.get("n_scales", len(key_tensor_field_kwargs["r_cluster_multiscale"]))
        self.time_emb_mlp = list(time_emb_mlp)
        self.time_mlps_multiscale = nn.ModuleList()          # parameters exist in the reference's state_dict; unused here
        for _ in range(self.n_scales):
            mods = []
            for i in range(1, len(time_emb_mlp)):
                mods.append(nn.Linear(time_emb_mlp[i - 1], time_emb_mlp[i]))
                if i != len(time_emb_mlp) - 1:
                    mods.append(nn.SiLU(inplace=True))
            self.time_mlps_multiscale.append(nn.Sequential(*mods))
        self.query_time_mlp = None
        self.time_emb_dim = time_emb_mlp[-1]
        # the reference mutates the kwargs dict in place (score_head_ebm.py:81-93); keep that contract
        assert "irreps_query" not in key_tensor_field_kwargs and "edge_context_emb_dim" not in key_tensor_field_kwargs
        key_tensor_field_kwargs["irreps_query"] = None
        key_tensor_field_kwargs["edge_context_emb_dim"] = None
        self.key_tensor_field = MultiscaleTensorField(**key_tensor_field_kwargs)
        self.irreps_key_edf = self.key_tensor_field.irreps_output
        self.key_edf_dim = self.irreps_key_edf.dim
        self.irreps_query_edf = Irreps(irreps_query_edf)
        self.query_edf_dim = self.irreps_query_edf.dim
        if self.irreps_query_edf != self.irreps_key_edf:
            raise NotImplementedError("query and key EDF irreps must agree")
        self.query_transform = TransformPcd(irreps=self.irreps_query_edf)
        self.register_buffer("q_indices", torch.tensor([[1, 2, 3], [0, 3, 2], [3, 0, 1], [2, 1, 0]], dtype=torch.long), persistent=False)
        self.register_buffer("q_factor", torch.tensor([[-0.5, -0.5, -0.5], [0.5, -0.5, 0.5], [0.5, 0.5, -0.5], [-0.5, 0.5, 0.5]]), persistent=False)
        self.energy_rescale_factor = 1.0 / float(self.key_edf_dim)
        self.inference_mode = False

    @torch.no_grad()
    def compute_energy(self, Ts: torch.Tensor, key_pcd_multiscale: List[FeaturedPoints], query_pcd: FeaturedPoints,
                       time: torch.Tensor) -> torch.Tensor:
        """energy[t] = sum_q w_q |field(T_t x_q) - D(q_t) psi_q|^2 / F   (score_head_ebm.py:122-174) -> (nT,)"""
        assert Ts.ndim == 2 and Ts.shape[-1] == 7, f"{Ts.shape}"
        assert time.ndim == 1 and len(time) == len(Ts), f"{time.shape}"
        assert query_pcd.f.ndim == 2 and query_pcd.f.shape[-1] == self.query_edf_dim, f"{query_pcd.f.shape}"
        assert isinstance(query_pcd.w, torch.Tensor)
        Ts = Ts.to(torch.float32).contiguous()
        nT, nQ = len(Ts), len(query_pcd.x)
        xq, fq = ops.query_transform(Ts, query_pcd.x.contiguous(), query_pcd.f.contiguous(), self.irreps_query_edf.m)
        bq = query_pcd.b.unsqueeze(0).expand(nT, -1).reshape(-1).contiguous()
        flat = FeaturedPoints(x=xq, f=fq, b=bq, w=None)
        field = self.key_tensor_field(query_points=flat, input_points_multiscale=key_pcd_multiscale)
        return ops.ebm_energy(field.f, fq, query_pcd.w.contiguous(), nT, nQ, self.energy_rescale_factor)

    def warmup(self, Ts, key_pcd_multiscale, query_pcd, time):
        return self.compute_energy(Ts=Ts, key_pcd_multiscale=key_pcd_multiscale, query_pcd=query_pcd, time=time)

    def train(self, mode: bool = True):
        super().train(mode=mode)
        self.inference_mode = not mode
        return self

    def forward(self, Ts, key_pcd_multiscale, query_pcd, time):
        """(ang_vel (nT, 3), lin_vel (nT, 3)) = pose gradient of log P = -energy in the body frame (score_head_ebm.py:192-222).
        Inference mode only (after ``.eval()``: ``create_graph=False`` in the reference); training the energy-based head
        through its score needs a double backward, which is not built."""
        assert Ts.ndim == 2 and Ts.shape[-1] == 7, f"{Ts.shape}"
        assert time.ndim == 1 and len(time) == len(Ts), f"{time.shape}"
        assert query_pcd.f.ndim == 2 and query_pcd.f.shape[-1] == self.query_edf_dim, f"{query_pcd.f.shape}"
        if not self.inference_mode:
            raise NotImplementedError("EbmScoreModelHead.forward in train mode (create_graph=True: a double backward through the "
                                      "field) is not built; call .eval() first")
        from . import train_path
        return train_path.ebm_score(self, Ts, key_pcd_multiscale, query_pcd)
