"""Equiformer blocks of the key encoder (UNet) and of the tensor field.

Mirrors /root/reference/diffusion_edf/block.py:64-174 (``EquiformerBlock`` used by the
UNet; note the reference computes ``norm_1_src`` / ``norm_1_dst`` and discards the
result, :149-153 -- the linears see un-normalised inputs; preserved here) and
/root/reference/diffusion_edf/gnn_block.py:65-218 (``EquiformerBlock`` of the score
head's ``MultiscaleTensorField``).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch import nn

from . import _lib as L
from . import ops
from .irreps import Irreps
from .layers import (EquivariantLayerNormV2, FeedForwardNetwork, GaussianRadialBasisLayerFiniteCutoff, GraphAttention,
                     LinearRS, ProjectIfMismatch)


def mlp_mid(irreps_emb: Irreps, mult) -> Irreps:
    if isinstance(mult, int):
        return irreps_emb * mult
    return Irreps(mult)


def node_tail(proj: LinearRS, norm: EquivariantLayerNormV2, ffn: FeedForwardNetwork, attn: torch.Tensor,
              res1: Optional[torch.Tensor]) -> torch.Tensor:
    """y1 = proj(attn) (+ res1);  y1 + ffn(norm(y1)): one dedf_node_chain launch when the irreps allow it, else three
    dedf_node_linear launches (same arithmetic, bit-identical)."""
    emb, pre = proj.irreps_out, ffn.fctp_1.irreps_out
    if proj.irreps_in == emb and ffn.fctp_1.irreps_in == emb and ffn.fctp_2.irreps_out == emb and ops.node_chain_ok(emb.m, pre.m):
        (P, pb), (A, ab), (B, bb) = proj.packed(), ffn.fctp_1.packed(), ffn.fctp_2.packed()
        return ops.node_chain(attn.contiguous(), emb.m, pre.m, P, pb, res1.contiguous() if res1 is not None else None,
                              norm.affine_weight.detach(), norm.affine_bias.detach(), norm.eps, A, ab, B, bb)
    out = proj(attn, res=res1) if res1 is not None else proj(attn)
    return ffn(out, ln=norm, res=out)


class UnetEquiformerBlock(nn.Module):
    def __init__(self, irreps_src, irreps_dst, irreps_edge_attr, irreps_head, num_heads: int, fc_neurons: Sequence[int],
                 irreps_mlp_mid=3, src_bias: bool = False, dst_bias: bool = True, **_ignored):
        super().__init__()
        self.irreps_src, self.irreps_dst = Irreps(irreps_src), Irreps(irreps_dst)
        self.irreps_emb = self.irreps_dst
        if Irreps(irreps_edge_attr).m != (1, 1, 1):
            raise NotImplementedError("edge attributes must be the l<=2 spherical harmonics 1x0e+1x1e+1x2e")
        self.norm_1_src = EquivariantLayerNormV2(self.irreps_src)      # parameters exist, output discarded (reference quirk)
        self.linear_src = LinearRS(self.irreps_src, self.irreps_emb, bias=src_bias)
        self.norm_1_dst = EquivariantLayerNormV2(self.irreps_dst)
        self.linear_dst = LinearRS(self.irreps_dst, self.irreps_emb, bias=dst_bias)
        self.ga = GraphAttention(self.irreps_emb, self.irreps_dst, fc_neurons, num_heads)
        self.norm_2 = EquivariantLayerNormV2(self.irreps_dst)
        self.ffn = FeedForwardNetwork(self.irreps_dst, self.irreps_dst, mlp_mid(self.irreps_emb, irreps_mlp_mid))

    def radial_weights(self, g: ops.Csr, length: torch.Tensor, radial: GaussianRadialBasisLayerFiniteCutoff) -> torch.Tensor:
        """Per-edge TP weights RadialProfile(GaussianRadialBasisLayerFiniteCutoff(length)): geometry only, no features
        (the key encoder computes them on a side stream while the previous block's features are still in flight)."""
        rad = self.ga.sep_act.dtp_rad
        E = max(1, g.n_edges)
        w = torch.empty(E, self.ga.sep_act.numel, dtype=torch.float32, device=length.device)
        d = L.MlpDesc()
        d.mode = L.MLP_IN_RBF
        d.n_edges_dev = L.ptr(g.n_edges_dev, torch.int32)
        d.length = L.ptr(length)
        mean, std, wl = radial.mean.detach().reshape(-1), radial.std_logit.detach().reshape(-1), radial.weight_logit.detach().reshape(-1)
        d.rbf_mean, d.rbf_std_logit, d.rbf_weight_logit = L.ptr(mean), L.ptr(std), L.ptr(wl)
        d.rbf_cutoff, d.rbf_offset = radial.cutoff, radial.offset
        rad.fill_desc(d, 0)
        d.out = L.ptr(w)
        if ops.USE_TC_MLP and d.W_tc[0]:
            ops.edge_mlp_tc(d, g.n_edges)
        else:
            ops.edge_mlp(d, g.n_edges)
        return w

    def forward(self, f_src: torch.Tensor, f_dst: torch.Tensor, g: ops.Csr, sh: torch.Tensor, length: torch.Tensor,
                radial: GaussianRadialBasisLayerFiniteCutoff, w: Optional[torch.Tensor] = None, w_ready=None) -> torch.Tensor:
        ms = list(self.irreps_src.m) + list(self.irreps_dst.m) + list(self.irreps_emb.m)
        if ops.USE_LINEAR_PAIR and all(m > 0 and m % 4 == 0 for m in ms):
            (Ws, bs), (Wd, bd) = self.linear_src.packed(), self.linear_dst.packed()
            msg_src, msg_dst = ops.node_linear_pair(f_src.contiguous(), self.irreps_src.m, Ws, bs, f_dst.contiguous(), self.irreps_dst.m, Wd, bd,
                                                    self.irreps_emb.m)
        else:
            msg_src = self.linear_src(f_src)
            msg_dst = self.linear_dst(f_dst)
        if w is None:
            w = self.radial_weights(g, length, radial)
        elif w_ready is not None:
            w_ready()        # ``w`` is produced on another stream: the caller's wait, placed after the node linears
        attn = self.ga.attend(msg_src, msg_dst, g, sh, w, None)
        return node_tail(self.ga.proj, self.norm_2, self.ffn, attn, f_dst)


class EquiformerBlock(nn.Module):
    """The tensor-field block (no destination features: ``use_dst_feature=False`` in every shipped config)."""

    def __init__(self, irreps_src, irreps_dst, irreps_edge_attr, num_heads: int, fc_neurons: Sequence[int], irreps_emb=None,
                 irreps_output=None, irreps_mlp_mid=3, use_dst_feature: bool = True, skip_connection: bool = True,
                 bias: bool = True, use_src_point_attn: bool = False, use_dst_point_attn: bool = False,
                 use_edge_weights: bool = True, **_ignored):
        super().__init__()
        if use_dst_feature or use_dst_point_attn or not use_edge_weights or not skip_connection:
            raise NotImplementedError("only the edge-time-encoding tensor field (no dst features / dst point attention) is "
                                      "implemented on the CUDA path")
        self.use_src_point_attn = use_src_point_attn
        self.irreps_src = Irreps(irreps_src)
        self.irreps_emb = Irreps(irreps_emb) if irreps_emb is not None else Irreps(irreps_dst)
        self.irreps_output = Irreps(irreps_output) if irreps_output is not None else Irreps(irreps_dst)
        self.prenorm_src = EquivariantLayerNormV2(self.irreps_src)
        self.linear_src = LinearRS(self.irreps_src, self.irreps_emb, bias=True)
        self.skip_2 = ProjectIfMismatch(self.irreps_emb, self.irreps_output, bias=True, layernorm=False)
        self.ga = GraphAttention(self.irreps_emb, self.irreps_emb, fc_neurons, num_heads, sh_lmax=Irreps(irreps_edge_attr).lmax)
        self.post_norm = EquivariantLayerNormV2(self.irreps_emb)
        self.ffn = FeedForwardNetwork(self.irreps_emb, self.irreps_output, mlp_mid(self.irreps_emb, irreps_mlp_mid))

    def source_messages(self, f_src: torch.Tensor) -> torch.Tensor:
        """linear_src(prenorm_src(f)): pose-independent, so callers may cache it per scene (gnn_block.py:170-171)."""
        return self.linear_src(f_src, ln=self.prenorm_src)

    def forward(self, msg_src: torch.Tensor, g: ops.Csr, sh: torch.Tensor, w: torch.Tensor,
                edge_logit: torch.Tensor, src_weight: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert (src_weight is not None) == self.use_src_point_attn, "source-point attention needs the source weights (FeaturedPoints.w)"
        attn = self.ga.attend(msg_src, None, g, sh, w, edge_logit, src_weight)
        if self.skip_2.is_identity:
            return node_tail(self.ga.proj, self.post_norm, self.ffn, attn, None)
        emb = self.ga.proj(attn)
        return self.ffn(emb, ln=self.post_norm, res=self.skip_2(emb))
