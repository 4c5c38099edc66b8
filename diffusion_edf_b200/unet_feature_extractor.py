"""Key-side encoder (4-scale equivariant U-Net) on the CUDA path.

Mirrors /root/reference/diffusion_edf/unet_feature_extractor.py:19-417 (constructor kwargs,
module tree / parameter names, forward semantics incl. the radius schedule quirk :79-86, the
(a+b)/sqrt(3) skips :347,359 and the skipped finest un-pool :383-384) with connectivity.py:8-76
(FpsPool / RadiusGraph) and utils.py:26-47 (ParityInversionSh).

Graph reuse (exact, set-wise): the up path runs on the down path's graphs with source and
destination swapped and the l=1 harmonics negated.  A radius graph is symmetric and
SH(-v) is the parity-flipped SH(v), so the swapped self-graph IS the original self-graph;
the swapped pool graph is the radius search with the roles of the two clouds exchanged.
Only the order in which a destination's edges are summed differs from the reference.
"""
from __future__ import annotations

import math
from typing import List, Optional, Union

import torch
from torch import nn

from . import ops
from .block import UnetEquiformerBlock
from .gnn_data import FeaturedPoints
from .irreps import Irreps
from .layers import GaussianRadialBasisLayerFiniteCutoff, LinearRS, ProjectIfMismatch


class ParityInversionSh(nn.Module):
    def __init__(self, irreps):
        super().__init__()
        m = Irreps(irreps).m
        self.register_buffer("sign", torch.cat([(1.0 if l % 2 == 0 else -1.0) * torch.ones((2 * l + 1) * m[l]) for l in range(3)]))


class _Geom:
    """A graph with its edge geometry (length, spherical harmonics)."""
    __slots__ = ("g", "length", "sh")

    def __init__(self, g, length, sh):
        self.g, self.length, self.sh = g, length, sh


class UnetFeatureExtractor(nn.Module):
    def __init__(self, irreps_input, irreps_output, irreps_emb: List, irreps_edge_attr: List, num_heads: List[int],
                 fc_neurons: List[List[int]], n_layers: List[int], pool_ratio: List[float], radius: List[Optional[float]],
                 deterministic: bool = False, pool_method="fps", irreps_mlp_mid=3, attn_type="mlp", alpha_drop=0.1,
                 proj_drop=0.1, drop_path_rate=0.0, n_layers_midstream: int = 2, n_scales: Optional[int] = None,
                 output_scalespace: Optional[List[int]] = None):
        super().__init__()
        self.irreps_output = Irreps(irreps_output)
        self.irreps_emb = [Irreps(i) for i in irreps_emb]
        self.irreps_edge_attr = [Irreps(i) for i in irreps_edge_attr]
        self.num_heads, self.fc_neurons, self.pool_ratio, self.n_layers = num_heads, fc_neurons, pool_ratio, n_layers
        self.deterministic = deterministic
        self.n_layers_midstream = n_layers_midstream
        if irreps_input is None:
            raise NotImplementedError("irreps_input=None")
        self.irreps_input = Irreps(irreps_input)
        self.input_emb = LinearRS(self.irreps_input, self.irreps_emb[0], bias=True)
        self.n_scales = len(self.irreps_emb) if n_scales is None else n_scales
        self.output_scalespace = list(range(self.n_scales)) if output_scalespace is None else \
            [self.n_scales + n if n < 0 else n for n in output_scalespace]
        self.radius = [radius[0]]
        for n, r in enumerate(radius[1:]):
            self.radius.append(self.radius[-1] / math.sqrt(self.pool_ratio[n - 1]) if r is None else r)
        pm = pool_method if isinstance(pool_method, list) else [pool_method] * self.n_scales
        if any(p != "fps" for p in pm) or any(n < 1 for n in n_layers):
            raise NotImplementedError("pool_method must be 'fps' and n_layers >= 1")
        mid = irreps_mlp_mid if isinstance(irreps_mlp_mid, list) else [irreps_mlp_mid] * self.n_scales
        at = attn_type if isinstance(attn_type, list) else [attn_type] * self.n_scales
        if any(a != "mlp" for a in at):
            raise NotImplementedError("attn_type must be 'mlp'")
        head = [self.irreps_emb[n].div(num_heads[n]) for n in range(self.n_scales)]

        def layer(n, src, dst, head_irreps):
            return nn.ModuleDict({
                "radial": GaussianRadialBasisLayerFiniteCutoff(num_basis=fc_neurons[n][0], cutoff=0.99 * self.radius[n]),
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

    @staticmethod
    def _run(layer, f_src, f_dst, geom: _Geom):
        return layer["gnn"](f_src, f_dst, geom.g, geom.sh, geom.length, layer["radial"])

    @staticmethod
    def _geom(x_src, x_dst, g) -> _Geom:
        length, sh, _ = ops.edge_geom(x_src, x_dst, g)
        return _Geom(g, length, sh)

    def forward(self, pcd: FeaturedPoints) -> List[FeaturedPoints]:
        x, b = pcd.x.contiguous(), pcd.b.contiguous()
        assert pcd.f.ndim == 2 and x.ndim == 2 and b.ndim == 1 and len(pcd.f) == len(x) == len(b)
        f = self.input_emb(pcd.f.contiguous())
        outs, graphs = [(f, x, b)], []          # graphs: ("pool", n, idx, x_src, b_src) / ("self", geom)
        geom = None
        for n, blk in enumerate(self.down_blocks):
            # ---- FPS pooling + bipartite radius graph (connectivity.py:59-76) ----
            idx = ops.fps(x, b, self.pool_ratio[n], random_start=not self.deterministic)
            x_dst = ops.gather_rows(x, idx)
            b_dst = b.index_select(0, idx)
            f_dst = ops.gather_rows(f, idx)
            g = ops.radius_csr(x, x_dst, [self.radius[n]], b_src=b, b_dst=b_dst, excl_mode=1, excl=idx, max_num_neighbors=1000)
            f_dst = blk["pool_proj"](f_dst)
            f_new = self._run(blk["pool_layer"], f, f_dst, self._geom(x, x_dst, g))
            graphs.append(("pool", n, idx, x, b))
            f, x, b = f_new, x_dst, b_dst
            outs.append((f, x, b))
            # ---- self radius graph + remaining layers ----
            g = ops.radius_csr(x, x, [self.radius[n]], b_src=b, b_dst=b, excl_mode=2, max_num_neighbors=1001)
            geom = self._geom(x, x, g)
            for layer in blk["layer_stack"]:
                f = self._run(layer, f, f, geom)
                outs.append((f, x, b))
                graphs.append(("self", geom))
        for layer in self.mid_block:
            f = self._run(layer, f, f, geom)
        f_skip, _, _ = outs.pop()
        f = ops.add_scale(f, f_skip, 1.0 / math.sqrt(3))
        ups = []
        for n, blk in enumerate(self.up_blocks):
            for layer in blk["layer_stack"]:
                f_dst, x_dst, b_dst = outs.pop()
                kind = graphs.pop()
                assert kind[0] == "self"
                f_dst = ops.add_scale(f, f_dst, 1.0 / math.sqrt(3))
                f = self._run(layer, f, f_dst, kind[1])      # swapped self graph == the same graph (see module docstring)
                x, b = x_dst, b_dst
            ups.append((f, x, b))
            f_dst, x_dst, b_dst = outs.pop()
            kind = graphs.pop()
            assert kind[0] == "pool"
            if n != self.n_scales - 1:
                _, scale, idx, x_fine, b_fine = kind
                # swapped pool graph: sources = pooled points (current x), destinations = finer points
                g = ops.radius_csr(x, x_fine, [self.radius[scale]], b_src=b, b_dst=b_fine, excl_mode=3, excl=idx,
                                   max_num_neighbors=1000)
                f = self._run(blk["unpool_layer"], f, f_dst, self._geom(x, x_fine, g))
                x, b = x_dst, b_dst
        ups = ups[::-1]
        pcds = []
        for s, proj in enumerate(self.project_outputs):
            if s not in self.output_scalespace:
                continue
            fs, xs, bs = ups[s]
            pcds.append(FeaturedPoints(x=xs, f=proj(fs), b=bs, w=None))
        return pcds
