"""Parameter containers that mirror the reference's module tree (so that
``state_dict`` keys agree with checkpoints trained by the reference, SURVEY.md
App. C) and run their arithmetic through the CUDA kernels in libdedf.so.

Reference classes mirrored (under /root/reference/diffusion_edf):
  equiformer/tensor_product_rescale.py:176-185 LinearRS, :155-173 / :241-268 FCTP (+SwishGate)
  equiformer/layer_norm.py:64-156 EquivariantLayerNormV2
  equiformer/radial_func.py:11-59 RadialProfile
  equiformer/graph_attention_transformer.py:60-135 SeparableFCTP
  graph_attention.py:16-122 GraphAttentionMLP, :138-273 GraphAttentionMLP2
  gnn_block.py:21-57 / block.py:21-57 FeedForwardNetwork
  skip.py:13-35 ProjectIfMismatch
  radial_func.py:168-227 GaussianRadialBasis, :231-278 GaussianRadialBasisLayerFiniteCutoff
None of the ``forward`` methods falls back to PyTorch arithmetic.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import _lib as L
from . import ops
from .irreps import Irreps, dtp_numel, dtp_out, dtp_paths, gate_pre, is_fused_family


def _versions(mod: nn.Module):
    return tuple((p.data_ptr(), p._version) for p in mod.parameters())


class _Packed:
    """Cache of kernel-layout copies of a module's parameters, rebuilt when they change."""

    def __init__(self):
        self._key = None
        self._val = None

    def get(self, mod: nn.Module, build):
        key = _versions(mod)
        if key != self._key:
            with torch.no_grad():
                self._val = build()
            self._key = key
        return self._val


def pack_tc(W: torch.Tensor, f16: bool = False) -> torch.Tensor:
    """(K, N) "in x out" fp32 weights -> the tensor-core kernel's B-operand stream (include/dedf.h, dedf_mlp_desc.W_tc):
    tf32 hi / lo split, N blocks of <= 256 columns, 8-wide K chunks, each part chunk-major [2][Nb][4].
    ``f16``: fp16 hi / lo split (hi = fp16(w), lo = fp16(w - hi)), 16-wide K chunks, each part [2][Nb][8 halves] -- byte for byte
    the same chunk shape; returned as a float32 VIEW of the half buffer."""
    K, N = W.shape
    nb = (N + 255) // 256
    assert K % 8 == 0 and N % nb == 0
    nbw = N // nb
    W = W.detach().contiguous().float()
    if f16:
        assert K % 16 == 0
        hi = W.half()
        parts = torch.stack([hi, (W - hi.float()).half()])           # (2, K, N)
        v = parts.view(2, K // 16, 2, 8, nb, nbw)                    # [part, kc, j, r, nb, n]
        return v.permute(4, 1, 0, 2, 5, 3).contiguous().view(-1).view(torch.float32)
    hi = (W.view(torch.int32) & -8192).view(torch.float32)          # clear the 13 low mantissa bits
    parts = torch.stack([hi, W - hi])                                # (2, K, N)
    v = parts.view(2, K // 8, 2, 4, nb, nbw)                         # [part, kc, j, r, nb, n]
    return v.permute(4, 1, 0, 2, 5, 3).contiguous().view(-1)         # [nb, kc, part, j, n, r]


def pack_tp_act_tc(G: int, W0: torch.Tensor, W1: torch.Tensor, W2: torch.Tensor, f16: bool = False) -> torch.Tensor:
    """Block-diagonal weights of the attention block's per-edge linear layer -> the B-operand stream of dedf_edge_tp_act_tc.

    W0 (D0, N0) = [sep_alpha | sep_act.lin 0e], W1 (D1, m1), W2 (D2, m2), rows in the depthwise tensor product's i_out order
    (SURVEY App. E).  The kernel consumes K in chunks of input channels: chunk j = 0e channels 8j..8j+7, 1e channels
    4j..4j+3, 2e channels 2j..2j+1; producer warp i (0..3) owns 0e channels a = 8j+2i, b = a+1 and 1e channel p = 4j+i,
    producer warp 4+t owns 2e channel q = 2j+t.  Per chunk and GEMM the K columns are grouped in fours (one float4 store of
    one producer thread), pads are zero rows:
        l_out=0 (16): group i = [k0 a, k0 b, k4 p, k12 q_i if i < 2 else 0]
        l_out=1 (24): group i = [k3 p, k5 p, k7 p, k1 a];  group 4 = [k1 b_0..b_3];  group 5 = [k10 q0, k13 q0, k10 q1, k13 q1]
        l_out=2 (24): group i = [k2 a, k2 b, k6 p, k8 p];  group 4+t = [k9 q_t, k11 q_t, k14 q_t, 0]
    Each block is stored chunk-major [K/4][N padded to 16][4], the chunk as [l0 | l1+l2 | l2] hi then the same for lo; the
    middle block holds the l_out=1 and l_out=2 weights side by side (the m = 4 rows of the 2e output share the 1e MMA).

    ``f16``: the fp16 hi / lo pack of the kind::f16 kernel variant -- chunks 2jj and 2jj+1 share one stage, the 16-byte K-group of a
    row holding the four values of the even chunk then the four of the odd one as halves ([K/4][N][8] halves: byte-for-byte the
    shape of a tf32 chunk), hi = fp16(w), lo = fp16(w - hi); returned as a float32 VIEW of the half buffer."""
    M0, M1, M2 = 2 * G, G, G // 2
    C0 = dict(k0=0, k4=M0, k12=M0 + M1)
    C1 = dict(k1=0, k3=M0, k5=M0 + M1, k7=M0 + 2 * M1, k10=M0 + 3 * M1, k13=M0 + 3 * M1 + M2)
    C2 = dict(k2=0, k6=M0, k8=M0 + M1, k9=M0 + 2 * M1, k11=M0 + 2 * M1 + M2, k14=M0 + 2 * M1 + 2 * M2)
    dev = W0.device
    W0, W1, W2 = (w.detach().float().cpu().contiguous() for w in (W0, W1, W2))     # a few hundred tiny ops: do them on the host
    assert W0.shape[0] == M0 + M1 + M2 and W1.shape == (M0 + 3 * M1 + 2 * M2, M1) and W2.shape == (M0 + 2 * M1 + 3 * M2, M2)

    def rows_of(W, rows):
        # (K, N) gather of weight rows, zero rows for the pads -- one index_select instead of a copy per row
        idx = torch.tensor([r if r is not None else W.shape[0] for r in rows], dtype=torch.long, device=W.device)
        return torch.cat([W, W.new_zeros(1, W.shape[1])])[idx]

    def block(W, rows):
        K, N = len(rows), W.shape[1]
        Np = (N + 15) // 16 * 16
        sel = torch.zeros(K, Np, dtype=torch.float32, device=W.device)
        sel[:, :N] = rows_of(W, rows)
        return sel.view(K // 4, 4, Np).permute(0, 2, 1).contiguous().view(-1)      # [K/4][Np][4]

    def block2(Wa, ra, Wb, rb):
        # [Wa | Wb] side by side: the 1e rows and the 2e m = 4 rows share one MMA (rows 0..95 read the Wa columns, 96..127 the Wb ones)
        K, Na, Nb = len(ra), Wa.shape[1], Wb.shape[1]
        Nap, Nbp = (Na + 15) // 16 * 16, (Nb + 15) // 16 * 16
        sel = torch.zeros(K, Nap + Nbp, dtype=torch.float32, device=Wa.device)
        sel[:, :Na] = rows_of(Wa, ra)
        sel[:, Nap:Nap + Nb] = rows_of(Wb, rb)
        return sel.view(K // 4, 4, Nap + Nbp).permute(0, 2, 1).contiguous().view(-1)

    chunks = []
    for j in range(M0 // 8):
        a = [8 * j + 2 * i for i in range(4)]
        p = [4 * j + i for i in range(4)]
        q = [2 * j, 2 * j + 1]
        r0, r1, r2 = [], [], []
        for i in range(4):
            r0 += [C0["k0"] + a[i], C0["k0"] + a[i] + 1, C0["k4"] + p[i], C0["k12"] + q[i] if i < 2 else None]
            r1 += [C1["k3"] + p[i], C1["k5"] + p[i], C1["k7"] + p[i], C1["k1"] + a[i]]
            r2 += [C2["k2"] + a[i], C2["k2"] + a[i] + 1, C2["k6"] + p[i], C2["k8"] + p[i]]
        r1 += [C1["k1"] + a[i] + 1 for i in range(4)]
        r1 += [C1["k10"] + q[0], C1["k13"] + q[0], C1["k10"] + q[1], C1["k13"] + q[1]]
        for t in range(2):
            r2 += [C2["k9"] + q[t], C2["k11"] + q[t], C2["k14"] + q[t], None]
        full = torch.cat([block(W0, r0), block2(W1, r1, W2, r2), block(W2, r2)])
        if f16:
            chunks.append(full)
            continue
        hi = (full.view(torch.int32) & -8192).view(torch.float32)
        chunks += [hi, full - hi]
    if f16:
        parts = []
        for jj in range(len(chunks) // 2):
            pair = torch.cat([chunks[2 * jj].view(-1, 4), chunks[2 * jj + 1].view(-1, 4)], dim=1)        # (K-groups x N, 8)
            hi = pair.half()
            parts += [hi.reshape(-1), (pair - hi.float()).half().reshape(-1)]
        return torch.cat(parts).contiguous().view(torch.float32).to(dev)
    return torch.cat(chunks).contiguous().to(dev)


def tp_act_w_perm(G: int) -> torch.Tensor:
    """Column order of the per-edge tensor-product weights that dedf_edge_tp_act_tc reads with vector loads (w_perm = 1):
    new column c holds original column perm[c].  Chunk j (60 columns): for producer warp i = 0..3 the 12 weights
    [k0 a, k0 b, k1 a, k1 b, k2 a, k2 b, k3 p .. k8 p] (a = 8j+2i, b = a+1, p = 4j+i), then for t = 0, 1 the 6 weights
    [k9 q .. k14 q] (q = 2j+t)."""
    M0, M1, M2 = 2 * G, G, G // 2
    perm = []
    for j in range(M0 // 8):
        for i in range(4):
            a, p = 8 * j + 2 * i, 4 * j + i
            perm += [a, a + 1, M0 + a, M0 + a + 1, 2 * M0 + a, 2 * M0 + a + 1] + [3 * M0 + s * M1 + p for s in range(6)]
        for t in range(2):
            q = 2 * j + t
            perm += [3 * M0 + 6 * M1 + s * M2 + q for s in range(6)]
    assert sorted(perm) == list(range(3 * M0 + 6 * M1 + 6 * M2))
    return torch.tensor(perm, dtype=torch.long)


def tc_mlp_ok(dims: Sequence[int]) -> bool:
    """Can dedf_edge_mlp_tc run an MLP with these layer widths?"""
    n = len(dims) - 1
    if n < 1 or n > L.MLP_MAX_LAYERS:
        return False
    for i in range(n):
        K, N = dims[i], dims[i + 1]
        nb = (N + 255) // 256
        if K < 8 or K % 8 or K > 128 or N < 16 or N % nb or (N // nb) % 16 or N > 512:
            return False
        if i < n - 1 and not (N == 128 or N <= 64):
            return False
    return True


class _TP(nn.Module):
    """Stands in for e3nn's o3.TensorProduct: only owns ``weight`` (key ``tp.weight``)."""

    def __init__(self, numel: int, internal: bool):
        super().__init__()
        if internal and numel > 0:
            self.weight = nn.Parameter(torch.randn(numel))
        else:
            self.register_parameter("weight", None)


class LinearRS(nn.Module):
    """Block-diagonal (per-l) linear map; also FullyConnectedTensorProductRescale with a 1x0e second operand."""

    def __init__(self, irreps_in, irreps_out, bias: bool = True, rescale: bool = True):
        super().__init__()
        self.irreps_in, self.irreps_out = Irreps(irreps_in), Irreps(irreps_out)
        shapes = [(a, b) for a, b in zip(self.irreps_in.m, self.irreps_out.m)]
        self._shapes = [(a, b) if (a and b) else None for a, b in shapes]
        numel = sum(a * b for s in self._shapes if s for a, b in [s])
        self.tp = _TP(numel, True)
        self.bias = nn.ParameterList([nn.Parameter(torch.zeros(self.irreps_out.m[0]))] if (bias and self.irreps_out.m[0]) else [])
        if rescale and numel:
            with torch.no_grad():      # tensor_product_rescale.py:97-120: weights ~ N(0,1) / sqrt(fan_in)
                off = 0
                for s in self._shapes:
                    if s:
                        self.tp.weight[off:off + s[0] * s[1]].mul_(1.0 / math.sqrt(s[0]))
                        off += s[0] * s[1]
        self._packed = _Packed()

    def packed(self):
        def build():
            w, off, out = self.tp.weight, 0, []
            for s in self._shapes:
                if s:
                    out.append(w[off:off + s[0] * s[1]].detach().clone().contiguous())
                    off += s[0] * s[1]
                else:
                    out.append(None)
            b = self.bias[0].detach().contiguous() if len(self.bias) else None
            return out, b
        return self._packed.get(self, build)

    def forward(self, x: torch.Tensor, ln: Optional["EquivariantLayerNormV2"] = None, gate: bool = False,
                res: Optional[torch.Tensor] = None, res_scale: float = 1.0) -> torch.Tensor:
        W, b = self.packed()
        lnp = None
        if ln is not None:
            lnp = (ln.affine_weight.detach(), ln.affine_bias.detach())
        return ops.node_linear(x.contiguous(), self.irreps_in.m, self.irreps_out.m, W, b, ln=lnp,
                               ln_eps=ln.eps if ln is not None else 1e-5, gate=gate, res=res, res_scale=res_scale)


class EquivariantLayerNormV2(nn.Module):
    """Parameters only; the normalisation runs as the prologue of the following LinearRS kernel."""

    def __init__(self, irreps, eps: float = 1e-5, affine: bool = True):
        super().__init__()
        self.irreps, self.eps = Irreps(irreps), eps
        assert affine
        self.affine_weight = nn.Parameter(torch.ones(self.irreps.num_irreps))
        self.affine_bias = nn.Parameter(torch.zeros(self.irreps.m[0]))


class ProjectIfMismatch(nn.Module):
    def __init__(self, irreps_in, irreps_out, bias: bool = True, layernorm: bool = True):
        super().__init__()
        self.irreps_in, self.irreps_out = Irreps(irreps_in), Irreps(irreps_out)
        if self.irreps_in == self.irreps_out:
            self.skip, self.layernorm = nn.Identity(), nn.Identity()
        else:
            self.layernorm = EquivariantLayerNormV2(self.irreps_in) if layernorm else nn.Identity()
            self.skip = LinearRS(self.irreps_in, self.irreps_out, bias=bias, rescale=True)

    @property
    def is_identity(self) -> bool:
        return isinstance(self.skip, nn.Identity)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.is_identity:
            return x
        ln = self.layernorm if isinstance(self.layernorm, EquivariantLayerNormV2) else None
        return self.skip(x, ln=ln)


class FeedForwardNetwork(nn.Module):
    def __init__(self, irreps_in, irreps_out, irreps_mid):
        super().__init__()
        self.irreps_mid = Irreps(irreps_mid)
        self.fctp_1 = LinearRS(irreps_in, gate_pre(self.irreps_mid))     # ...SwishGate: linear to the pre-gate irreps
        self.fctp_2 = LinearRS(self.irreps_mid, irreps_out)

    def forward(self, x: torch.Tensor, ln: EquivariantLayerNormV2, res: torch.Tensor) -> torch.Tensor:
        """res + fctp_2(Gate(fctp_1(LN(x))))"""
        h = self.fctp_1(x, ln=ln, gate=True)
        return self.fctp_2(h, res=res, res_scale=1.0)


class RadialProfile(nn.Module):
    def __init__(self, ch_list: Sequence[int]):
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
        self.ch_list = list(ch_list)
        self._packed = _Packed()
        self._packed_perm = _Packed()
        # set by the consumer (GraphAttention): column order in which dedf_edge_tp_act_tc wants the output (tp_act_w_perm)
        self.out_perm: Optional[torch.Tensor] = None

    def permuted(self) -> bool:
        """Is the kernel-side output in ``out_perm`` column order right now?  (The parameters / state_dict never are.)"""
        return self.out_perm is not None and ops.USE_TC_TPACT

    def packed(self):
        perm = self.out_perm if self.permuted() else None

        def build():
            lin = [m for m in self.net if isinstance(m, nn.Linear)]
            lns = [m for m in self.net if isinstance(m, nn.LayerNorm)]
            W = [m.weight.detach().t().contiguous() for m in lin]
            b = [m.bias.detach().contiguous() if m.bias is not None else None for m in lin]
            g = [m.weight.detach().contiguous() for m in lns]
            bb = [m.bias.detach().contiguous() for m in lns]
            off = self.offset.detach().contiguous()
            if perm is not None:      # a column permutation of the output = of the last layer's weights, bias and offset
                pm = perm.to(W[-1].device)
                W[-1] = W[-1][:, pm].contiguous()
                b[-1] = b[-1][pm].contiguous() if b[-1] is not None else None
                off = off[pm].contiguous()
            Wtc = [pack_tc(w) for w in W] if tc_mlp_ok(self.ch_list) else None
            Wtc16 = [pack_tc(w, f16=True) for w in W] if Wtc is not None and all(k % 16 == 0 for k in self.ch_list[:-1]) else None
            return W, b, g, bb, off, Wtc, Wtc16
        return (self._packed_perm if perm is not None else self._packed).get(self, build)

    def f16_ok(self) -> bool:
        """fp16 hi / lo split on the tensor cores (ops.MLP_F16) possible for these layer widths?"""
        return ops.MLP_F16 and self.packed()[6] is not None

    def fill_desc(self, d: L.MlpDesc, first_layer: int = 0, f16: Optional[bool] = None) -> None:
        """Describe the MLP layers starting at slot ``first_layer`` of ``d`` (dims[first_layer] must be ch_list[0]).
        ``f16``: hand over the fp16 packs (None: whenever possible; a caller that adds layers of its own in front decides)."""
        W, b, g, bb, off, Wtc, Wtc16 = self.packed()
        if f16 is None:
            f16 = self.f16_ok()
        if f16:
            Wtc = Wtc16
        d.tc_f16 = 1 if f16 else 0
        n = len(W)
        assert first_layer + n <= L.MLP_MAX_LAYERS
        for i in range(n):
            s = first_layer + i
            d.dims[s] = self.ch_list[i]
            d.dims[s + 1] = self.ch_list[i + 1]
            d.W[s] = L.ptr(W[i])
            d.W_tc[s] = L.ptr(Wtc[i]) if Wtc is not None else None
            d.b[s] = L.ptr(b[i])
            last = i == n - 1
            d.ln_g[s] = None if last else L.ptr(g[i])
            d.ln_b[s] = None if last else L.ptr(bb[i])
            d.flags[s] = 0 if last else 3
        d.n_layers = first_layer + n
        d.out_offset = L.ptr(off)
        # keep the packed tensors alive for as long as the descriptor is
        d._keep = (W, b, g, bb, off, Wtc)


class GaussianRadialBasisLayerFiniteCutoff(nn.Module):
    """Parameters of the UNet's edge-length encoder; evaluated inside dedf_edge_mlp (RBF input mode)."""

    def __init__(self, num_basis: int, cutoff: float):
        super().__init__()
        self.num_basis, self.cutoff = num_basis, float(cutoff)
        self.offset = 0.01 * self.cutoff
        self.mean = nn.Parameter(torch.linspace(0, 1.0, num_basis + 2)[1:-1].unsqueeze(0))
        self.std_logit = nn.Parameter(torch.full((1, num_basis), math.log(math.exp(2.0 / num_basis) - 1)))
        self.weight_logit = nn.Parameter(torch.full((1, num_basis), -math.log(4.0 / 1.0 - 1)))


class _GaussianParamModule(nn.Module):
    def __init__(self, dim: int, max_weight: float = 4.0):
        super().__init__()
        self.std_logit = nn.Parameter(torch.full((1, dim), math.log(math.exp(2.0 / dim) - 1), dtype=torch.float32))
        self.weight_logit = nn.Parameter(torch.full((1, dim), -math.log(max_weight / 1.0 - 1), dtype=torch.float32))
        self.mean = nn.Parameter(torch.linspace(0.0, 1.0, dim + 2, dtype=torch.float32)[1:-1].unsqueeze(0))


class GaussianRadialBasis(nn.Module):
    def __init__(self, dim: int, max_val: float):
        super().__init__()
        self.dim, self.max_val = int(dim), float(max_val)
        self.param_module = _GaussianParamModule(dim)


class _DTP(nn.Module):
    """DepthwiseTensorProduct container: ``tp.weight`` exists only for shared (internal) weights."""

    def __init__(self, numel: int, internal: bool, fan_in_scales: Optional[torch.Tensor] = None):
        super().__init__()
        self.tp = _TP(numel, internal)
        if internal and fan_in_scales is not None:
            with torch.no_grad():
                self.tp.weight.mul_(fan_in_scales)


class SeparableFCTP(nn.Module):
    """dtp (+ dtp_rad) + lin (+ gate).  Node irreps ``m0=2G, m1=G, m2=G/2`` with the l<=2 harmonics (every shipped config) run on
    the fused edge kernels; any other even-parity l<=2 irreps / harmonics degree keeps the same parameters (same state_dict keys)
    and runs un-fused on the table-driven depthwise tensor product (``fused_family`` False, ``paths``)."""

    def __init__(self, irreps_node: Irreps, irreps_out: Irreps, fc_neurons: Optional[Sequence[int]], use_activation: bool,
                 internal_weights: bool, sh_lmax: int = 2):
        super().__init__()
        self.irreps_node, self.irreps_out = Irreps(irreps_node), Irreps(irreps_out)
        self.sh_lmax = int(sh_lmax)
        self.fused_family = is_fused_family(self.irreps_node) and self.sh_lmax == 2 and all(self.irreps_out.m)
        # output filter of DepthwiseTensorProduct = the l's of the node OUTPUT irreps (0e always kept)
        filt = None if self.fused_family else self.irreps_out
        self.paths = dtp_paths(self.irreps_node, self.sh_lmax, (0, 1, 2) if filt is None else tuple(l for l in range(3) if filt.m[l]))
        self.irreps_dtp_out = dtp_out(self.irreps_node, self.sh_lmax, filt)
        self.numel = dtp_numel(self.irreps_node, self.sh_lmax, filt)
        self.dtp = _DTP(self.numel, internal_weights)
        self.dtp_rad = RadialProfile(list(fc_neurons) + [self.numel]) if fc_neurons is not None else None
        self.use_activation = use_activation
        lin_out = gate_pre(self.irreps_out) if use_activation else self.irreps_out
        self.lin = LinearRS(self.irreps_dtp_out, lin_out)


class GraphAttention(nn.Module):
    """GraphAttentionMLP / GraphAttentionMLP2 (same parameters; the latter adds the edge logits)."""

    def __init__(self, irreps_emb, irreps_out, fc_neurons: Sequence[int], num_heads: int, sh_lmax: int = 2):
        super().__init__()
        self.irreps_emb, self.irreps_out = Irreps(irreps_emb), Irreps(irreps_out)
        if num_heads != 4:
            raise NotImplementedError("the attention kernels are specialised for 4 heads (all shipped configs)")
        self.num_heads = num_heads
        self.irreps_head = self.irreps_emb.div(num_heads)
        mul_alpha = self.irreps_emb.m[0]
        self.sep_act = SeparableFCTP(self.irreps_emb, self.irreps_emb, fc_neurons, use_activation=True, internal_weights=False, sh_lmax=sh_lmax)
        self.fused_family = self.sep_act.fused_family
        self.sep_alpha = LinearRS(Irreps((self.sep_act.irreps_dtp_out.m[0], 0, 0)), Irreps((mul_alpha, 0, 0)))
        self.sep_value = SeparableFCTP(self.irreps_emb, self.irreps_emb, None, use_activation=False, internal_weights=True, sh_lmax=sh_lmax)
        self.alpha_dot = nn.Parameter(torch.randn(1, num_heads, mul_alpha // num_heads))
        nn.init.xavier_uniform_(self.alpha_dot)
        self.proj = LinearRS(self.irreps_emb, self.irreps_out)
        self._packed = _Packed()
        if self.irreps_emb.m[1] in (16, 32) and self.irreps_emb.m == (2 * self.irreps_emb.m[1], self.irreps_emb.m[1], self.irreps_emb.m[1] // 2):
            self.sep_act.dtp_rad.out_perm = tp_act_w_perm(self.irreps_emb.m[1])

    def packed(self):
        def build():
            (a0, _, _), ab = self.sep_alpha.packed()
            (l0, l1, l2), lb = self.sep_act.lin.packed()
            d0 = dtp_out(self.irreps_emb).m[0]
            ma = self.irreps_emb.m[0]
            n_lin0 = self.sep_act.lin.irreps_out.m[0]
            W0 = torch.cat([a0.view(d0, ma), l0.view(d0, n_lin0)], dim=1).contiguous()
            b0 = torch.cat([ab, lb]).contiguous()
            (v0, v1, v2), vb = self.sep_value.lin.packed()
            G = self.irreps_emb.m[1]
            d1, d2 = dtp_out(self.irreps_emb).m[1], dtp_out(self.irreps_emb).m[2]
            Wtc = pack_tp_act_tc(G, W0, l1.view(d1, -1), l2.view(d2, -1)) if G in (16, 32) else None
            Wtc16 = pack_tp_act_tc(G, W0, l1.view(d1, -1), l2.view(d2, -1), f16=True) if G in (16, 32) else None
            return dict(W0=W0, W1=l1, W2=l2, Wtc=Wtc, Wtc16=Wtc16, b0=b0, alpha_dot=self.alpha_dot.detach().reshape(-1).contiguous(),
                        wv=self.sep_value.dtp.tp.weight.detach().contiguous(), V0=v0, V1=v1, V2=v2, vb=vb)
        return self._packed.get(self, build)

    def attend(self, msg_src: torch.Tensor, msg_dst: Optional[torch.Tensor], g: ops.Csr, sh: torch.Tensor,
               w: torch.Tensor, edge_logit: Optional[torch.Tensor], src_weight: Optional[torch.Tensor] = None,
               w_perm: Optional[bool] = None) -> torch.Tensor:
        """-> sum_e softmax(logit)_e * value_e per destination, (n_dst, F) (before ``proj``).  ``src_weight`` (N_src,): source-point
        attention, alpha_e *= w[src_e] AFTER the softmax (graph_attention.py:258-259) == scaling the value rows.
        ``w``: per-edge tensor-product weights as produced by the kernels from ``self.sep_act.dtp_rad.packed()`` (column order
        ``dtp_rad.permuted()``); pass ``w_perm=False`` for weights in the reference's order."""
        if not self.fused_family:
            raise L.DedfError("irreps outside the fused kernels' family run through train_path.graph_attention (un-fused kernels)")
        p = self.packed()
        G = self.irreps_emb.m[1]
        F = self.irreps_emb.dim
        E = max(1, g.n_edges)
        dev = msg_src.device
        logits = torch.empty(E, 4, dtype=torch.float32, device=dev)
        v = torch.empty(E, F, dtype=torch.float32, device=dev)
        if ops.USE_TC_TPACT and p["Wtc"] is not None:
            # ``w`` comes from self.sep_act.dtp_rad, whose kernel-side output is in the chunk-major column order while the flag is on
            f16 = ops.TPACT_F16 and p.get("Wtc16") is not None
            ops.edge_tp_act_tc(G, msg_src, msg_dst, g, sh, w, self.sep_act.numel, p["Wtc16"] if f16 else p["Wtc"], p["b0"], p["alpha_dot"],
                               edge_logit, logits, v, w_perm=self.sep_act.dtp_rad.permuted() if w_perm is None else w_perm, f16=f16)
        else:
            if self.sep_act.dtp_rad.permuted() if w_perm is None else w_perm:
                raise L.DedfError("tensor-product weights in chunk-major column order need dedf_edge_tp_act_tc")
            ops.edge_tp_lin(G, L.EPI_ACT, msg_src, msg_dst, False, g, sh, w, self.sep_act.numel, p["W0"], p["W1"], p["W2"], p["b0"],
                            alpha_dot=p["alpha_dot"], edge_logit=edge_logit, logits=logits, out=v)
        if ops.USE_VALUE_REDUCE:
            # reassociated value path: reduce the tensor-product outputs over the incoming edges, THEN the linear layer
            post = ops.edge_gather_scalar(src_weight, g) if src_weight is not None else None
            return ops.value_reduce(G, g, v, sh, logits, post, p["wv"], p["V0"], p["V1"], p["V2"], p["vb"])
        val = torch.empty(E, F, dtype=torch.float32, device=dev)
        ops.edge_tp_lin(G, L.EPI_LIN, v, None, True, g, sh, p["wv"], 0, p["V0"], p["V1"], p["V2"], p["vb"], out=val)
        if src_weight is not None:
            val = ops.row_scale(val, ops.edge_gather_scalar(src_weight, g), self.irreps_emb.m)
        return ops.segment_softmax_reduce(g, logits, val, self.irreps_emb.m)
