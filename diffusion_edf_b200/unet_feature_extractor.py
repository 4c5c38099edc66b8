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

import contextlib
import math
import os
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
                 output_scalespace: Optional[List[int]] = None, forward_only: bool = False):
        super().__init__()
        self.forward_only = forward_only         # ForwardOnlyFeatureExtractor: down path only
        self.irreps_output = Irreps(irreps_output)
        self.irreps_emb = [Irreps(i) for i in irreps_emb]
        self.irreps_edge_attr = [Irreps(i) for i in irreps_edge_attr]
        self.num_heads, self.fc_neurons, self.pool_ratio, self.n_layers = num_heads, fc_neurons, pool_ratio, n_layers
        self.deterministic = deterministic
        self.alpha_drop, self.proj_drop = float(alpha_drop), float(proj_drop)     # train mode only (train_path.py)
        if drop_path_rate and float(drop_path_rate) > 0.0:
            raise NotImplementedError("drop_path_rate > 0 (GraphDropPath, block.py:133,163-171) is not implemented; every shipped config uses 0.0")
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
                                        for _ in range(0 if forward_only else n_layers_midstream)])
        self.up_blocks = nn.ModuleList()
        for n in (() if forward_only else range(self.n_scales - 1, -1, -1)):
            blk = nn.ModuleDict()
            blk["parity_inversion"] = ParityInversionSh(self.irreps_edge_attr[n])
            blk["layer_stack"] = nn.ModuleList([layer(n, self.irreps_emb[n], self.irreps_emb[n], head[n])
                                                for _ in range(n_layers[n] - 1)])
            blk["unpool_layer"] = layer(n, self.irreps_emb[n], self.irreps_emb[max(n - 1, 0)], head[max(n - 1, 0)])
            self.up_blocks.append(blk)
        self.project_outputs = nn.ModuleList([ProjectIfMismatch(self.irreps_emb[n], self.irreps_output)
                                              for n in range(self.n_scales)])
        self.overlap_geometry = True     # graph construction + radial MLPs on a side stream (see _geometry)
        self._side = None
        self._fps_stream = None

    @staticmethod
    def _geom(x_src, x_dst, g) -> _Geom:
        length, sh, _ = ops.edge_geom(x_src, x_dst, g)
        return _Geom(g, length, sh)

    # ------------------------------------------------------------------ geometry pass (feature independent)
    def _geometry(self, x: torch.Tensor, b: torch.Tensor, done, fps_stream=None):
        """FPS pooling, radius graphs, edge geometry and the per-edge radial TP weights of EVERY block, in the order the
        feature pass consumes them.  Nothing here reads a feature, so forward() runs it on a side stream: after the first
        FPS the whole graph-construction / radial-MLP pipeline overlaps the attention blocks of the previous scales.
        ``done(item)`` is called after each item's kernels are enqueued (it records the event the feature pass waits on)."""
        items = []

        def emit(kind, **kw):
            it = dict(kind=kind, **kw)
            done(it)
            items.append(it)
            return it

        def block(layer, geom):
            return layer["gnn"].radial_weights(geom.g, geom.length, layer["radial"])

        levels = []          # per scale: (x_src, b_src, idx, x_dst, b_dst, self geom)
        geom = None
        # The FPS chain of ALL scales first, on its own stream when there is one: level n + 1 only needs the coordinates level n
        # picked, not its graphs / radial MLPs, and an FPS launch occupies 1-8 SMs.  In the in-graph timeline
        # (profiles/r2_s8_timeline_128_before.txt) FPS 1-3 sat in the middle of the geometry chain and the main stream idled for them.
        pooled = []
        cur = torch.cuda.current_stream() if x.is_cuda else None
        if fps_stream is not None:
            fps_stream.wait_stream(cur)
        xx, bb = x, b
        for n in range(len(self.down_blocks)):
            with torch.cuda.stream(fps_stream) if fps_stream is not None else contextlib.nullcontext():
                idx = ops.fps(xx, bb, self.pool_ratio[n], random_start=not self.deterministic)
                ops.stamp(f"geo fps{n}")
                x_dst = ops.gather_rows(xx, idx)
                b_dst = bb.index_select(0, idx)
                ev = None
                if fps_stream is not None:
                    ev = torch.cuda.Event()
                    ev.record(fps_stream)
                    for t in (idx, x_dst, b_dst):
                        t.record_stream(cur)
            pooled.append((idx, x_dst, b_dst, ev))
            xx, bb = x_dst, b_dst
        for n, blk in enumerate(self.down_blocks):
            idx, x_dst, b_dst, ev = pooled[n]
            if ev is not None:
                cur.wait_event(ev)
            g = ops.radius_csr(x, x_dst, [self.radius[n]], b_src=b, b_dst=b_dst, excl_mode=1, excl=idx, max_num_neighbors=1000)
            gp = self._geom(x, x_dst, g)
            ops.stamp(f"geo pool-graph{n}")
            ev_g = None
            if fps_stream is not None:       # the block's node linears only need the graph: they start under the radial MLP
                ev_g = torch.cuda.Event()
                ev_g.record(cur)
            emit("pool", n=n, idx=idx, x_dst=x_dst, b_dst=b_dst, geom=gp, w=block(blk["pool_layer"], gp), event_g=ev_g)
            ops.stamp(f"geo pool-mlp{n}")
            g = ops.radius_csr(x_dst, x_dst, [self.radius[n]], b_src=b_dst, b_dst=b_dst, excl_mode=2, max_num_neighbors=1001)
            geom = self._geom(x_dst, x_dst, g)
            ops.stamp(f"geo self-graph{n}")
            for layer in blk["layer_stack"]:
                emit("self", geom=geom, w=block(layer, geom))
            ops.stamp(f"geo self-mlps{n}")
            levels.append((x, b, idx, x_dst, b_dst, geom))
            x, b = x_dst, b_dst
        for layer in self.mid_block:
            emit("self", geom=geom, w=block(layer, geom))
        for n, blk in enumerate(self.up_blocks):                     # (empty for the forward-only encoder)
            scale = self.n_scales - 1 - n
            x_fine, b_fine, idx, x_c, b_c, gself = levels[scale]
            for layer in blk["layer_stack"]:
                emit("self", geom=gself, w=block(layer, gself))      # swapped self graph == the same graph
            if n != self.n_scales - 1:
                # swapped pool graph: sources = pooled points, destinations = finer points
                g = ops.radius_csr(x_c, x_fine, [self.radius[scale]], b_src=b_c, b_dst=b_fine, excl_mode=3, excl=idx,
                                   max_num_neighbors=1000)
                gu = self._geom(x_c, x_fine, g)
                emit("unpool", geom=gu, w=block(blk["unpool_layer"], gu))
                ops.stamp(f"geo unpool{scale}")
        return items

    def forward(self, pcd: FeaturedPoints) -> List[FeaturedPoints]:
        x, b = pcd.x.contiguous(), pcd.b.contiguous()
        assert pcd.f.ndim == 2 and x.ndim == 2 and b.ndim == 1 and len(pcd.f) == len(x) == len(b)
        main = torch.cuda.current_stream()
        use_side = self.overlap_geometry and x.is_cuda
        if use_side:
            if self._side is None or self._side.device != x.device:
                # high priority: the geometry chain is many small kernels the feature pass waits on; without it they queue behind the
                # feature pass's chip-filling kernels (in-graph timeline: the main stream idled ~0.3 ms per forward for them)
                self._side = torch.cuda.Stream(device=x.device, priority=-1 if os.environ.get("DEDF_GEOM_PRIO", "1") != "0" else 0)
            side = self._side
            side.wait_stream(main)                       # fork (inputs are ready on the main stream)

            def done(it):
                ev = torch.cuda.Event()
                ev.record(side)
                it["event"] = ev
            if self._fps_stream is None or self._fps_stream.device != x.device:
                self._fps_stream = torch.cuda.Stream(device=x.device, priority=-1 if os.environ.get("DEDF_GEOM_PRIO", "1") != "0" else 0)
            with torch.cuda.stream(side):
                items = self._geometry(x, b, done, self._fps_stream)
        else:
            items = self._geometry(x, b, lambda it: None)
        pos = [0]

        def take(kind):
            it = items[pos[0]]
            pos[0] += 1
            assert it["kind"] == kind, (it["kind"], kind)
            if use_side:
                main.wait_event(it["event_g"] if it.get("event_g") is not None else it["event"])
                for v in it.values():                    # tensors allocated on the side stream, consumed on the main one
                    for t in (v if isinstance(v, (tuple, list)) else (v,)):
                        if isinstance(t, torch.Tensor):
                            t.record_stream(main)
                gm = it["geom"]
                for t in (gm.length, gm.sh, gm.g.row_ptr, gm.g.edge_src, gm.g.edge_dst):
                    t.record_stream(main)
            return it

        def run(layer, f_src, f_dst, it):
            gm = it["geom"]
            ops.stamp(f"blk start {it['kind']} n_dst={gm.g.n_dst}")
            w_ready = (lambda: main.wait_event(it["event"])) if use_side and it.get("event_g") is not None else None
            out = layer["gnn"](f_src, f_dst, gm.g, gm.sh, gm.length, layer["radial"], w=it["w"], w_ready=w_ready)
            ops.stamp(f"blk end {it['kind']} n_dst={gm.g.n_dst}")
            return out

        f = self.input_emb(pcd.f.contiguous())
        outs = [(f, x, b)]
        scale_outs = []
        for n, blk in enumerate(self.down_blocks):
            it = take("pool")
            f_dst = blk["pool_proj"](ops.gather_rows(f, it["idx"]))
            f = run(blk["pool_layer"], f, f_dst, it)
            x, b = it["x_dst"], it["b_dst"]
            outs.append((f, x, b))
            for layer in blk["layer_stack"]:
                f = run(layer, f, f, take("self"))
                outs.append((f, x, b))
            scale_outs.append((f, x, b))
        if self.forward_only:            # forward_only_feature_extractor.py:191-275: every scale outputs its down-path features
            assert pos[0] == len(items)
            if use_side:
                main.wait_stream(side)
            return [FeaturedPoints(x=scale_outs[s][1], f=proj(scale_outs[s][0]), b=scale_outs[s][2], w=None)
                    for s, proj in enumerate(self.project_outputs) if s in self.output_scalespace]
        for layer in self.mid_block:
            f = run(layer, f, f, take("self"))
        f_skip, _, _ = outs.pop()
        f = ops.add_scale(f, f_skip, 1.0 / math.sqrt(3))
        ups = []
        for n, blk in enumerate(self.up_blocks):
            for layer in blk["layer_stack"]:
                f_dst, x_dst, b_dst = outs.pop()
                f_dst = ops.add_scale(f, f_dst, 1.0 / math.sqrt(3))
                f = run(layer, f, f_dst, take("self"))
                x, b = x_dst, b_dst
            ups.append((f, x, b))
            f_dst, x_dst, b_dst = outs.pop()
            if n != self.n_scales - 1:
                f = run(blk["unpool_layer"], f, f_dst, take("unpool"))
                x, b = x_dst, b_dst
        assert pos[0] == len(items)
        if use_side:
            main.wait_stream(side)                       # join (required under CUDA-graph capture)
        ups = ups[::-1]
        pcds = []
        for s, proj in enumerate(self.project_outputs):
            if s not in self.output_scalespace:
                continue
            fs, xs, bs = ups[s]
            pcds.append(FeaturedPoints(x=xs, f=proj(fs), b=bs, w=None))
        return pcds


class ForwardOnlyFeatureExtractor(UnetFeatureExtractor):
    """Key encoder of the sapien* highres configs (/root/reference/diffusion_edf/forward_only_feature_extractor.py:19-275):
    the UNet's down path only; same constructor kwargs and the same parameter names as the UNet's down path."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs, forward_only=True)
