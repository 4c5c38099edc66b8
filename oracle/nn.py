"""Equivariant building blocks, restated on plain torch (CPU).  Oracle = test
infrastructure only; never imported by the product package.

Follows (file:line relative to /root/reference/diffusion_edf):
  equiformer/tensor_product_rescale.py:20-152 (TensorProductRescale), :155-173
  (FullyConnectedTensorProductRescale), :176-185 (LinearRS), :241-268
  (..SwishGate), :352-382 (DepthwiseTensorProduct); equiformer/fast_activation.py
  :14-23 (SmoothLeakyReLU), :31-152 (Activation), :156-224 (Gate);
  equiformer/layer_norm.py:64-156 (EquivariantLayerNormV2);
  equiformer/radial_func.py:11-59 (RadialProfile);
  equiformer/graph_attention_transformer.py:60-135 (SeparableFCTP), :139-201
  (Vec2AttnHeads / AttnHeads2Vec); skip.py:13-35 (ProjectIfMismatch).
e3nn==0.4.4 semantics per SURVEY.md App. A.4 / A.5.
"""
from __future__ import annotations

import functools
import math
from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

from .irreps import Irreps, irreps2gate, selection_rule, sort_even_first
from .so3 import wigner_3j


# --------------------------------------------------------------------------
# e3nn.math.normalize2mom  (App. A.5)
# --------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def _normal_sample() -> torch.Tensor:
    gen = torch.Generator(device="cpu").manual_seed(0)
    return torch.randn(1_000_000, generator=gen, dtype=torch.float64)


def normalize2mom_const(f) -> float:
    with torch.no_grad():
        cst = f(_normal_sample()).pow(2).mean().pow(-0.5).item()
    return 1.0 if abs(cst - 1.0) < 1e-4 else cst


def smooth_leaky_relu(x: torch.Tensor, alpha: float = 0.2) -> torch.Tensor:
    return ((1 + alpha) / 2) * x + ((1 - alpha) / 2) * x * (2 * torch.sigmoid(x) - 1)


@functools.lru_cache(maxsize=None)
def act_consts() -> dict:
    return {
        "silu": normalize2mom_const(torch.nn.functional.silu),
        "sigmoid": normalize2mom_const(torch.sigmoid),
        "slrelu": normalize2mom_const(smooth_leaky_relu),
    }


# --------------------------------------------------------------------------
# o3.TensorProduct(path_normalization='none', 'component')  (App. A.4)
# --------------------------------------------------------------------------
Instr = Tuple[int, int, int, str]  # (i_in1, i_in2, i_out, mode)


class TensorProduct(nn.Module):
    """``out = sum_paths sqrt(2 l_out + 1) * w . C . x1 . x2`` (uvu / uvw only)."""

    def __init__(self, in1: Irreps, in2: Irreps, out: Irreps, instructions: Sequence[Instr],
                 internal_weights: bool):
        super().__init__()
        self.in1, self.in2, self.out = Irreps(in1), Irreps(in2), Irreps(out)
        self.instructions = [tuple(i[:4]) for i in instructions]
        self.shapes = []
        for i1, i2, io, mode in self.instructions:
            m1, m2, mo = self.in1[i1][0], self.in2[i2][0], self.out[io][0]
            if mode == "uvu":
                assert m1 == mo
                self.shapes.append((m1, m2))
            elif mode == "uvw":
                self.shapes.append((m1, m2, mo))
            else:
                raise NotImplementedError(mode)
        self.weight_numel = sum(math.prod(s) for s in self.shapes)
        self.internal_weights = internal_weights
        if internal_weights and self.weight_numel > 0:
            self.weight = nn.Parameter(torch.randn(self.weight_numel))  # e3nn init: N(0,1)
        else:
            self.weight = None

    def weight_views(self, weight: torch.Tensor):
        off = 0
        for s in self.shapes:
            n = math.prod(s)
            yield weight[..., off:off + n].reshape(weight.shape[:-1] + s)
            off += n

    def forward(self, x1: torch.Tensor, x2: torch.Tensor, weight: Optional[torch.Tensor] = None):
        if weight is None:
            weight = self.weight
        assert weight is not None and weight.shape[-1] == self.weight_numel
        per_sample = weight.dim() == 2
        N = x1.shape[0]
        s1, s2 = self.in1.slices(), self.in2.slices()
        outs = [x1.new_zeros(N, m, 2 * l + 1) for m, l, _ in self.out]
        for (i1, i2, io, mode), w in zip(self.instructions, self.weight_views(weight)):
            (m1, l1, _), (m2, l2, _), (mo, lo, _) = self.in1[i1], self.in2[i2], self.out[io]
            a = x1[:, s1[i1]].reshape(N, m1, 2 * l1 + 1)
            b = x2[:, s2[i2]].reshape(N, m2, 2 * l2 + 1)
            C = wigner_3j(l1, l2, lo).to(device=x1.device, dtype=x1.dtype) * math.sqrt(2 * lo + 1)
            t = torch.einsum("zui,zvj,ijk->zuvk", a, b, C)
            z = "z" if per_sample else ""
            if mode == "uvu":
                outs[io] = outs[io] + torch.einsum(f"{z}uv,zuvk->zuk", w, t)
            else:
                outs[io] = outs[io] + torch.einsum(f"{z}uvw,zuvk->zwk", w, t)
        return torch.cat([o.reshape(N, m * (2 * l + 1)) for o, (m, l, _) in zip(outs, self.out)], dim=1)


class TensorProductRescale(nn.Module):
    """tensor_product_rescale.py:20-152.  ``rescale`` only touches the INITIAL
    weights (:103-120); one bias vector per 0e entry of the simplified output."""

    def __init__(self, in1, in2, out, instructions, bias=True, rescale=True, internal_weights=False):
        super().__init__()
        self.irreps_in1, self.irreps_in2, self.irreps_out = Irreps(in1), Irreps(in2), Irreps(out)
        self.tp = TensorProduct(self.irreps_in1, self.irreps_in2, self.irreps_out, instructions, internal_weights)
        simp = self.irreps_out.simplify()
        self.bias_slices: List[Tuple[int, int]] = []
        biases = []
        if bias:
            for (m, l, p), sl in zip(simp, simp.slices()):
                if l == 0 and p == 1:
                    biases.append(nn.Parameter(torch.zeros(m)))
                    self.bias_slices.append((sl.start, sl.stop))
        self.bias = nn.ParameterList(biases)
        # fan-in per output slice (:52-63, :97-120)
        fan = {}
        for (i1, i2, io, mode) in self.tp.instructions:
            f = {"uvw": self.irreps_in1[i1][0] * self.irreps_in2[i2][0], "uvu": self.irreps_in2[i2][0]}[mode]
            fan[io] = fan.get(io, 0) + f
        self.slices_sqrt_k = {}
        oslices = self.irreps_out.slices()
        for (i1, i2, io, mode) in self.tp.instructions:
            self.slices_sqrt_k[io] = (oslices[io], (1 / fan[io] ** 0.5) if rescale else 1.0)
        if internal_weights and rescale:
            with torch.no_grad():
                for w, (i1, i2, io, mode) in zip(self.tp.weight_views(self.tp.weight.data), self.tp.instructions):
                    w.mul_(1 / fan[io] ** 0.5)

    def forward(self, x, y, weight=None):
        out = self.tp(x, y, weight)
        for (a, b), bias in zip(self.bias_slices, self.bias):
            out = torch.cat([out[:, :a], out[:, a:b] + bias, out[:, b:]], dim=1)
        return out


def fctp_instructions(in1: Irreps, in2: Irreps, out: Irreps):
    return [
        (i1, i2, io, "uvw")
        for i1, (_, l1, p1) in enumerate(in1)
        for i2, (_, l2, p2) in enumerate(in2)
        for io, (_, lo, po) in enumerate(out)
        if (lo, po) in set(selection_rule(l1, p1, l2, p2))
    ]


class FullyConnectedTensorProductRescale(TensorProductRescale):
    def __init__(self, in1, in2, out, bias=True, rescale=True, internal_weights=True):
        in1, in2, out = Irreps(in1), Irreps(in2), Irreps(out)
        super().__init__(in1, in2, out, fctp_instructions(in1, in2, out), bias=bias, rescale=rescale,
                         internal_weights=internal_weights)


class LinearRS(FullyConnectedTensorProductRescale):
    def __init__(self, irreps_in, irreps_out, bias=True, rescale=True):
        super().__init__(Irreps(irreps_in), Irreps("1x0e"), Irreps(irreps_out), bias=bias, rescale=rescale)

    def forward(self, x):  # noqa: D401
        return super().forward(x, torch.ones_like(x[:, 0:1]))


class Gate(nn.Module):
    """fast_activation.py:156-224 with SiLU scalars / sigmoid gates, both
    multiplied by their normalize2mom constants."""

    def __init__(self, scalars: Irreps, gates: Irreps, gated: Irreps):
        super().__init__()
        self.scalars, self.gates, self.gated = Irreps(scalars), Irreps(gates), Irreps(gated)
        assert self.gates.num_irreps == self.gated.num_irreps
        self.irreps_in = (self.scalars + self.gates + self.gated).simplify()
        self.irreps_out = self.scalars + self.gated
        c = act_consts()
        self.c_silu, self.c_sig = c["silu"], c["sigmoid"]

    def forward(self, x):
        ns, ng = self.scalars.dim, self.gates.dim
        s = self.c_silu * torch.nn.functional.silu(x[..., :ns])
        if ng == 0:
            return s
        g = self.c_sig * torch.sigmoid(x[..., ns:ns + ng])
        parts, off, ig = [s], ns + ng, 0
        for m, l, _ in self.gated:
            d = 2 * l + 1
            blk = x[..., off:off + m * d].reshape(x.shape[:-1] + (m, d))
            parts.append((blk * g[..., ig:ig + m, None]).reshape(x.shape[:-1] + (m * d,)))
            off += m * d
            ig += m
        return torch.cat(parts, dim=-1)


class ScalarActivation(nn.Module):
    """fast_activation.py:31-152 for an all-scalar irreps with SiLU."""

    def __init__(self):
        super().__init__()
        self.c = act_consts()["silu"]

    def forward(self, x):
        return self.c * torch.nn.functional.silu(x)


def make_gate(irreps_out: Irreps):
    scal, gates, gated = irreps2gate(irreps_out)
    if gated.num_irreps == 0:
        return ScalarActivation(), Irreps(irreps_out)
    g = Gate(scal, gates, gated)
    return g, g.irreps_in


class FullyConnectedTensorProductRescaleSwishGate(FullyConnectedTensorProductRescale):
    def __init__(self, in1, in2, out, bias=True, rescale=True):
        gate, irreps_pre = make_gate(Irreps(out))
        super().__init__(in1, in2, irreps_pre, bias=bias, rescale=rescale)
        self.gate = gate

    def forward(self, x, y, weight=None):
        return self.gate(super().forward(x, y, weight))


def DepthwiseTensorProduct(irreps_in: Irreps, irreps_edge: Irreps, irreps_out_filter: Irreps,
                           internal_weights=False, bias=True, rescale=True) -> TensorProductRescale:
    """tensor_product_rescale.py:352-382."""
    irreps_in, irreps_edge, filt = Irreps(irreps_in), Irreps(irreps_edge), Irreps(irreps_out_filter)
    allowed = {(l, p) for _, l, p in filt}
    out, instr = [], []
    for i, (m, l1, p1) in enumerate(irreps_in):
        for j, (_, l2, p2) in enumerate(irreps_edge):
            for (lo, po) in selection_rule(l1, p1, l2, p2):
                if (lo, po) in allowed or (lo, po) == (0, 1):
                    instr.append((i, j, len(out), "uvu"))
                    out.append((m, lo, po))
    sorted_out, perm, _ = sort_even_first(Irreps(out))
    instr = [(i1, i2, perm[io], mode) for i1, i2, io, mode in instr]
    return TensorProductRescale(irreps_in, irreps_edge, sorted_out, instr, bias=bias, rescale=rescale,
                                internal_weights=internal_weights)


class EquivariantLayerNormV2(nn.Module):
    """equiformer/layer_norm.py:64-156 ('component')."""

    def __init__(self, irreps, eps=1e-5, affine=True):
        super().__init__()
        self.irreps, self.eps, self.affine = Irreps(irreps), eps, affine
        if affine:
            self.affine_weight = nn.Parameter(torch.ones(self.irreps.num_irreps))
            self.affine_bias = nn.Parameter(torch.zeros(self.irreps.count(0, 1)))
        else:
            self.register_parameter("affine_weight", None)
            self.register_parameter("affine_bias", None)

    def forward(self, x, batch=None):
        out, ix, iw, ib = [], 0, 0, 0
        for m, l, p in self.irreps:
            d = 2 * l + 1
            f = x[:, ix:ix + m * d].reshape(-1, m, d)
            ix += m * d
            if l == 0 and p == 1:
                f = f - f.mean(dim=1, keepdim=True)
            nrm = f.pow(2).mean(-1).mean(dim=1, keepdim=True)
            nrm = (nrm + self.eps).pow(-0.5)
            if self.affine:
                nrm = nrm * self.affine_weight[None, iw:iw + m]
                iw += m
            f = f * nrm.reshape(-1, m, 1)
            if self.affine and d == 1 and p == 1:
                f = f + self.affine_bias[ib:ib + m].reshape(m, 1)
                ib += m
            out.append(f.reshape(-1, m * d))
        return torch.cat(out, dim=-1)


class RadialProfile(nn.Module):
    """equiformer/radial_func.py:11-59."""

    def __init__(self, ch_list: List[int]):
        super().__init__()
        mods, cin = [], ch_list[0]
        for i in range(1, len(ch_list)):
            last = i == len(ch_list) - 1
            mods.append(nn.Linear(cin, ch_list[i], bias=not last))
            cin = ch_list[i]
            if last:
                break
            mods.append(nn.LayerNorm(ch_list[i]))
            mods.append(nn.SiLU())
        self.net = nn.Sequential(*mods)
        self.offset = nn.Parameter(torch.zeros(ch_list[-1]))
        bound = 1 / math.sqrt(ch_list[-2])
        nn.init.uniform_(self.offset, -bound, bound)

    def forward(self, x):
        return self.net(x) + self.offset.reshape(1, -1)


class SeparableFCTP(nn.Module):
    """equiformer/graph_attention_transformer.py:60-135 (norm_layer=None)."""

    def __init__(self, irreps_node_input, irreps_edge_attr, irreps_node_output, fc_neurons,
                 use_activation=False, internal_weights=False):
        super().__init__()
        self.irreps_node_input = Irreps(irreps_node_input)
        self.irreps_edge_attr = Irreps(irreps_edge_attr)
        self.irreps_node_output = Irreps(irreps_node_output)
        self.dtp = DepthwiseTensorProduct(self.irreps_node_input, self.irreps_edge_attr, self.irreps_node_output,
                                          bias=False, internal_weights=internal_weights)
        self.dtp_rad = None
        if fc_neurons is not None:
            self.dtp_rad = RadialProfile(list(fc_neurons) + [self.dtp.tp.weight_numel])
            with torch.no_grad():
                # :91-93 -- slices are OUTPUT slices applied to the weight rows, as in the reference
                for (sl, k) in self.dtp.slices_sqrt_k.values():
                    self.dtp_rad.net[-1].weight.data[sl, :] *= k
                    self.dtp_rad.offset.data[sl] *= k
        lin_out = self.irreps_node_output
        scal, gates, gated = irreps2gate(self.irreps_node_output)
        if use_activation:
            lin_out = (scal + gates + gated).simplify()
        self.lin = LinearRS(self.dtp.irreps_out.simplify(), lin_out)
        self.gate = None
        if use_activation:
            self.gate = ScalarActivation() if gated.num_irreps == 0 else Gate(scal, gates, gated)

    def forward(self, node_input, edge_attr, edge_scalars=None):
        weight = None
        if self.dtp_rad is not None and edge_scalars is not None:
            weight = self.dtp_rad(edge_scalars)
        out = self.lin(self.dtp(node_input, edge_attr, weight))
        if self.gate is not None:
            out = self.gate(out)
        return out


def vec2heads(x: torch.Tensor, irreps_head: Irreps, num_heads: int) -> torch.Tensor:
    """graph_attention_transformer.py:139-168."""
    N, out, off = x.shape[0], [], 0
    for m, l, _ in irreps_head:
        w = m * num_heads * (2 * l + 1)
        out.append(x[:, off:off + w].reshape(N, num_heads, m * (2 * l + 1)))
        off += w
    return torch.cat(out, dim=2)


def heads2vec(x: torch.Tensor, irreps_head: Irreps) -> torch.Tensor:
    """graph_attention_transformer.py:177-201."""
    N, out, off = x.shape[0], [], 0
    for m, l, _ in irreps_head:
        w = m * (2 * l + 1)
        out.append(x[:, :, off:off + w].reshape(N, x.shape[1] * w))
        off += w
    return torch.cat(out, dim=1)


class ProjectIfMismatch(nn.Module):
    """skip.py:13-35."""

    def __init__(self, irreps_in, irreps_out, bias=True, layernorm=True):
        super().__init__()
        self.irreps_in, self.irreps_out = Irreps(irreps_in), Irreps(irreps_out)
        if self.irreps_in == self.irreps_out:
            self.skip, self.layernorm = nn.Identity(), nn.Identity()
        else:
            self.layernorm = EquivariantLayerNormV2(self.irreps_in) if layernorm else nn.Identity()
            self.skip = LinearRS(self.irreps_in, self.irreps_out, bias=bias, rescale=True)

    def forward(self, x):
        return self.skip(self.layernorm(x))
