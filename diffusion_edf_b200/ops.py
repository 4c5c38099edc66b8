"""Thin Python wrappers over the C ABI (include/dedf.h).  PyTorch is used only to
own device memory and streams; every arithmetic op below is a hand-written
sm_100a kernel in libdedf.so.  All tensors must be CUDA / contiguous.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import List, NamedTuple, Optional, Sequence, Tuple

import torch

from . import _lib as L
from ._lib import check, ptr, stream


class Csr(NamedTuple):
    """Graph in CSR-by-destination form.  ``row_ptr`` has n_seg * n_dst + 1 entries."""
    row_ptr: torch.Tensor      # int32
    edge_src: torch.Tensor     # int32 (E)  flat source index
    edge_dst: torch.Tensor     # int32 (E)
    n_edges_dev: torch.Tensor  # int32 (1)  == row_ptr[-1:]
    n_edges: int
    n_dst: int
    n_seg: int


# ----------------------------------------------------------------------------
# call accounting: kernels launched (bench.py's ``gpu_launches``) and optional per-entry-point CUDA-event timing
# ----------------------------------------------------------------------------
GRID_MIN_SOURCES = 1024     # radius searches over at least this many sources use the grid-hash kernels
LAUNCHES = 0
PROFILE = None      # set to {} to collect {entry point: [(start_event, end_event), ...]}
FLOPS = {}          # while PROFILE is on: {entry point: algorithmic fp32 FLOPs (2 x MAC) of the calls made} -- bench.py's roofline_step


def _flops(name: str, value: float) -> None:
    if PROFILE is not None:
        FLOPS[name] = FLOPS.get(name, 0.0) + float(value)


def _irr_mac(a, b) -> int:
    """MACs per node of a block-diagonal LinearRS a -> b (each l block applied to its 2l+1 components)."""
    return a[0] * b[0] + 3 * a[1] * b[1] + 5 * a[2] * b[2]


def _tp_act_flops(G: int) -> float:
    """per edge: [sep_alpha | sep_act.lin] on the depthwise tensor-product output (104 448 at G = 32) + the CG contraction."""
    M0, M1, M2 = 2 * G, G, G // 2
    D0, D1, D2 = M0 + M1 + M2, M0 + 3 * M1 + 2 * M2, M0 + 2 * M1 + 3 * M2
    return 2.0 * (D0 * (M0 + D0) + 3 * D1 * M1 + 5 * D2 * M2) + 10_000.0 * G / 32


def _value_flops(G: int):
    """(per edge, per destination): CG contraction + 4 head-weighted folds of the 49 G outputs; sep_value.lin once per destination."""
    M0, M1, M2 = 2 * G, G, G // 2
    D0, D1, D2 = M0 + M1 + M2, M0 + 3 * M1 + 2 * M2, M0 + 2 * M1 + 3 * M2
    return 10_000.0 * G / 32 + 2.0 * 4 * 49 * G, 2.0 * (D0 * M0 + 3 * D1 * M1 + 5 * D2 * M2)
_KERNELS = {"dedf_grid_build": 4, "dedf_radius_grid_count": 2, "dedf_radius_grid_fill": 1, "dedf_fps": 1, "dedf_radius_count": 2, "dedf_radius_fill": 1, "dedf_edge_geom": 1, "dedf_edge_mlp": 1, "dedf_edge_mlp_tc": 1,
            "dedf_edge_tp_lin": 1, "dedf_segment_softmax_reduce": 1, "dedf_edge_tp_reduce": 1, "dedf_node_linear": 1,
            "dedf_gather_rows": 1, "dedf_weight_post": 1, "dedf_add_scale": 1, "dedf_time_embed": 1, "dedf_query_transform": 1, "dedf_score_tp": 1,
            "dedf_pose_update": 2, "dedf_sample_advance": 1, "dedf_prefetch_l2": 1, "dedf_tc_selftest": 1}


def _call(name: str, *args) -> None:
    global LAUNCHES
    fn = getattr(L.load(), name)
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        PROFILE.setdefault(name, []).append((e0, e1))
    else:
        rc = fn(*args)
    check(rc, name)
    LAUNCHES += _KERNELS.get(name, 1)


# ----------------------------------------------------------------------------
# static plans: make a data-dependent forward replayable as a CUDA graph
# ----------------------------------------------------------------------------
TIMELINE = None     # profiling only (profiles/run_timeline.py): {"buf": int64 device tensor, "names": []}


def stamp(name: str) -> None:
    """Record %globaltimer at this point of the current stream into the next slot of ``TIMELINE`` (no-op when it is None)."""
    tl = TIMELINE
    if tl is None:
        return
    i = len(tl["names"])
    if i >= tl["buf"].numel():
        return
    tl["names"].append((name, int(torch.cuda.current_stream().cuda_stream)))
    rc = L.load().dedf_stamp(tl["buf"].data_ptr() + 8 * i, stream())
    check(rc, "dedf_stamp")


class Plan:
    """Host-side values a forward pass normally has to read back from the device (edge counts, batch layout), recorded
    once in an eager pass and then replayed, so that the same Python code runs without any host synchronisation and with
    buffer sizes fixed ahead of time (edge counts become capacities; ``overflow`` is raised on the device if one is exceeded)."""

    def __init__(self, margin: float = 1.5, pad: int = 1024):
        self.items: list = []
        self.pos = 0
        self.mode = "record"
        self.margin, self.pad = margin, pad
        self.overflow: Optional[torch.Tensor] = None

    def start(self, mode: str, overflow: Optional[torch.Tensor] = None):
        self.mode, self.pos, self.overflow = mode, 0, overflow
        if mode == "record":
            self.items = []

    def value(self, fn):
        """Record ``fn()`` (eager pass) or return the recorded value (replay pass)."""
        if self.mode == "record":
            v = fn()
            self.items.append(v)
            return v
        v = self.items[self.pos]
        self.pos += 1
        return v

    def capacity(self, n_edges: int) -> int:
        return int(n_edges * self.margin) + self.pad


_PLAN: Optional[Plan] = None


class use_plan:
    def __init__(self, plan: Optional[Plan], mode: str = "record", overflow: Optional[torch.Tensor] = None):
        self.plan, self.mode, self.overflow = plan, mode, overflow

    def __enter__(self):
        global _PLAN
        self.prev = _PLAN
        _PLAN = self.plan
        if self.plan is not None:
            self.plan.start(self.mode, self.overflow)
        return self.plan

    def __exit__(self, *exc):
        global _PLAN
        _PLAN = self.prev
        return False


def plan_value(fn):
    """``fn()`` normally; under a plan the value is recorded / replayed (use for anything that syncs with the device)."""
    return fn() if _PLAN is None else _PLAN.value(fn)


def replaying() -> bool:
    return _PLAN is not None and _PLAN.mode == "replay"


# ----------------------------------------------------------------------------
# L2 weight prefetch
# ----------------------------------------------------------------------------
def prefetch_table(tensors: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor, int, list]:
    """Device table (pointers, byte counts) of the given CUDA tensors for ``prefetch_l2`` (keeps the tensors alive)."""
    ts = [t for t in tensors if t is not None and t.is_cuda and t.numel() > 0]
    dev = ts[0].device
    ptrs = torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64).to(dev)
    sizes = torch.tensor([t.numel() * t.element_size() for t in ts], dtype=torch.int64).to(dev)
    return ptrs, sizes, len(ts), ts


def prefetch_l2(table) -> None:
    ptrs, sizes, n, _ = table
    _call("dedf_prefetch_l2", ptrs.data_ptr(), sizes.data_ptr(), n, stream())


def flag_if_differs(a: torch.Tensor, b: torch.Tensor, flag: torch.Tensor) -> None:
    """flag |= 1 on the device if the int64 tensors differ (no host synchronisation)."""
    assert a.shape == b.shape
    _call("dedf_flag_if_differs", ptr(a, torch.long), ptr(b, torch.long), a.numel(), ptr(flag, torch.int32), stream())


def tc_selftest(A: torch.Tensor, B: torch.Tensor, n_split: int = 3) -> torch.Tensor:
    """D = A @ B.T on the tcgen05 tensor cores (A: (128, K), B: (N, K)); see include/dedf.h."""
    assert A.shape[0] == 128 and A.shape[1] == B.shape[1]
    D = torch.empty(128, B.shape[0], dtype=torch.float32, device=A.device)
    _call("dedf_tc_selftest", ptr(A.contiguous()), ptr(B.contiguous()), B.shape[0], A.shape[1], n_split, ptr(D), stream())
    return D


# ----------------------------------------------------------------------------
# graph construction
# ----------------------------------------------------------------------------
def fps(x: torch.Tensor, batch: Optional[torch.Tensor], ratio: float, random_start: bool = False) -> torch.Tensor:
    """torch_cluster.fps: LongTensor of selected indices, per batch segment, in selection order."""
    n_total = x.shape[0]
    x = x.contiguous()

    def _segments():
        if batch is None or n_total == 0:
            return [(0, n_total)]
        counts = torch.bincount(batch).tolist()        # batch ids are sorted/contiguous (as torch_cluster requires)
        segs, o = [], 0
        for c in counts:
            if c:
                segs.append((o, c))
            o += c
        return segs

    segs = plan_value(_segments)
    ms = [int(math.ceil(ratio * n)) for _, n in segs]
    out = torch.empty(sum(ms), dtype=torch.long, device=x.device)
    off = 0
    for (o, n), m in zip(segs, ms):
        # random start: drawn on the device and read by the kernel, so the call never synchronises
        start_dev = torch.randint(n, (1,), device=x.device, dtype=torch.long) if random_start else None
        scratch = torch.empty(n, dtype=torch.float32, device=x.device) if n > 16384 else None
        _call("dedf_fps", ptr(x) + o * 12, n, m, 0, ptr(start_dev, torch.long), o, out.data_ptr() + off * 8, ptr(scratch), stream())
        off += m
    return out


def radius_csr(x_src: torch.Tensor, x_dst: torch.Tensor, radii: Sequence[Optional[float]],
               src_off: Optional[Sequence[int]] = None, b_src: Optional[torch.Tensor] = None,
               b_dst: Optional[torch.Tensor] = None, excl_mode: int = 0, excl: Optional[torch.Tensor] = None,
               max_num_neighbors: int = 1000, capacity: Optional[int] = None,
               overflow: Optional[torch.Tensor] = None) -> Csr:
    """Radius search of ``x_dst`` against ``len(radii)`` concatenated source clouds (None radius = all pairs).

    Default: exact CSR, one host sync to size the edge buffers.  With ``capacity`` (or under a replaying ``Plan``) the edge
    buffers are pre-sized, nothing synchronises, the CSR is clamped on the device and ``overflow`` (int32[1]) is raised
    if the capacity was exceeded; ``Csr.n_edges`` is then the capacity and the true count lives in ``n_edges_dev``."""
    n_scales = len(radii)
    if src_off is None:
        assert n_scales == 1
        src_off = [0, x_src.shape[0]]
    n_dst = x_dst.shape[0]
    x_src, x_dst = x_src.contiguous(), x_dst.contiguous()
    so = L.int_array(src_off)
    rr = L.float_array([-1.0 if r is None else float(r) for r in radii])
    dev = x_src.device
    counts = torch.empty(max(1, n_scales * n_dst), dtype=torch.int32, device=dev)
    row_ptr = torch.empty(n_scales * n_dst + 1, dtype=torch.int32, device=dev)
    pb_s = ptr(b_src, torch.long) if b_src is not None else None
    pb_d = ptr(b_dst, torch.long) if b_dst is not None else None
    pex = ptr(excl, torch.long) if excl is not None else None
    if capacity is None and replaying():
        capacity = _PLAN.capacity(_PLAN.value(None))
        overflow = _PLAN.overflow
    # one large source cloud and a finite radius: grid-hash search (27 cells) instead of the ordered brute-force scan
    use_grid = (n_scales == 1 and radii[0] is not None and x_src.shape[0] >= GRID_MIN_SOURCES and n_dst > 0)
    if use_grid:
        n_src = x_src.shape[0]
        n_buckets = max(32, 1 << int(math.ceil(math.log2(2 * n_src))))
        bucket_cnt = torch.empty(n_buckets, dtype=torch.int32, device=dev)
        bucket_start = torch.empty(n_buckets + 1, dtype=torch.int32, device=dev)
        sorted_idx = torch.empty(n_src, dtype=torch.int32, device=dev)
        sorted_xyz = torch.empty(n_src, 3, dtype=torch.float32, device=dev)
        r = float(radii[0])
        _call("dedf_grid_build", ptr(x_src), n_src, r, n_buckets, ptr(bucket_cnt, torch.int32), ptr(bucket_start, torch.int32),
              ptr(sorted_idx, torch.int32), ptr(sorted_xyz), stream())
        stamp(f"  grid built n_src={n_src}")
        gargs = (ptr(x_src), n_src, ptr(x_dst), n_dst, r, n_buckets, ptr(bucket_start, torch.int32), ptr(sorted_idx, torch.int32),
                 ptr(sorted_xyz), pb_s, pb_d, excl_mode, pex, max_num_neighbors)
        cap = max(1, int(capacity)) if capacity is not None else 0
        _call("dedf_radius_grid_count", *gargs, ptr(counts, torch.int32), ptr(row_ptr, torch.int32), cap, None,
              ptr(overflow, torch.int32) if capacity is not None else None, stream())
        stamp(f"  grid counted n_dst={n_dst}")
        if capacity is None:
            n_edges_dev = row_ptr[-1:]
            n_edges = plan_value(lambda: int(n_edges_dev.item()))
            n_alloc = n_edges
        else:
            n_edges = n_alloc = cap
        edge_src = torch.empty(max(1, n_alloc), dtype=torch.int32, device=dev)
        edge_dst = torch.empty(max(1, n_alloc), dtype=torch.int32, device=dev)
        _call("dedf_radius_grid_fill", *gargs, ptr(row_ptr, torch.int32), ptr(edge_src, torch.int32), ptr(edge_dst, torch.int32), stream())
        stamp(f"  grid filled n_dst={n_dst}")
        return Csr(row_ptr, edge_src[:n_edges], edge_dst[:n_edges], row_ptr[-1:], n_edges, n_dst, 1)
    if capacity is not None:
        capacity = max(1, int(capacity))
        _call("dedf_radius_count", ptr(x_src), ptr(x_dst), n_dst, n_scales, so, rr, pb_s, pb_d, excl_mode, pex,
              max_num_neighbors, ptr(counts, torch.int32), ptr(row_ptr, torch.int32), capacity, None,
              ptr(overflow, torch.int32), stream())
        edge_src = torch.empty(capacity, dtype=torch.int32, device=dev)
        edge_dst = torch.empty(capacity, dtype=torch.int32, device=dev)
        _call("dedf_radius_fill", ptr(x_src), ptr(x_dst), n_dst, n_scales, so, rr, pb_s, pb_d, excl_mode, pex,
              max_num_neighbors, ptr(row_ptr, torch.int32), ptr(edge_src, torch.int32), ptr(edge_dst, torch.int32), stream())
        return Csr(row_ptr, edge_src, edge_dst, row_ptr[-1:], capacity, n_dst, n_scales)
    _call("dedf_radius_count", ptr(x_src), ptr(x_dst), n_dst, n_scales, so, rr, pb_s, pb_d, excl_mode, pex,
                                max_num_neighbors, ptr(counts, torch.int32), ptr(row_ptr, torch.int32), 0, None, None, stream())
    n_edges_dev = row_ptr[-1:]
    n_edges = plan_value(lambda: int(n_edges_dev.item()))   # the one host sync of a graph build (sizes the edge buffers)
    edge_src = torch.empty(max(1, n_edges), dtype=torch.int32, device=dev)
    edge_dst = torch.empty(max(1, n_edges), dtype=torch.int32, device=dev)
    _call("dedf_radius_fill", ptr(x_src), ptr(x_dst), n_dst, n_scales, so, rr, pb_s, pb_d, excl_mode, pex,
                               max_num_neighbors, ptr(row_ptr, torch.int32), ptr(edge_src, torch.int32),
                               ptr(edge_dst, torch.int32), stream())
    return Csr(row_ptr, edge_src[:n_edges], edge_dst[:n_edges], n_edges_dev, n_edges, n_dst, n_scales)


USE_HEAD_FRONT = True   # pose transform + radius search + CSR + edge geometry of the score head in one launch (dedf_head_front)
_FRONT_WS = {}          # (device, stream) -> (cta_sum, barrier): the kernel's device-wide barrier state, zeroed once


def head_front(Ts: torch.Tensor, qx: torch.Tensor, b_q: Optional[torch.Tensor], x_src: torch.Tensor, b_src: Optional[torch.Tensor],
               src_off: Sequence[int], radii: Sequence[Optional[float]], ns_cut: Tuple[float, float], max_num_neighbors: int = 1000,
               capacity: Optional[int] = None, overflow: Optional[torch.Tensor] = None, step: Optional[torch.Tensor] = None,
               rows_all: Optional[torch.Tensor] = None, rows_cur: Optional[torch.Tensor] = None, static_sources: bool = False):
    """-> (Csr, length (E), sh (E, 9), logit (E), x_dst (n_t * n_q, 3)); see include/dedf.h (dedf_head_front).  Exact CSR with one host
    sync to size the edge buffers, unless ``capacity`` is given (or a ``Plan`` is replaying): then nothing synchronises, the CSR is
    clamped on the device and ``overflow`` is raised if the capacity was exceeded."""
    n_t, n_q, n_scales = Ts.shape[0], qx.shape[0], len(radii)
    n_dst = n_t * n_q
    dev = Ts.device
    if capacity is None and replaying():
        capacity = _PLAN.capacity(_PLAN.value(None))
        overflow = _PLAN.overflow
    key = (dev, stream())
    if key not in _FRONT_WS:
        _FRONT_WS[key] = (torch.empty(256, dtype=torch.int32, device=dev), torch.zeros(2, dtype=torch.int32, device=dev))
    cta_sum, barrier = _FRONT_WS[key]
    d = L.HeadFrontDesc()
    d.Ts, d.n_t, d.qx, d.n_q = ptr(Ts), n_t, ptr(qx), n_q
    d.x_src, d.b_src, d.b_q = ptr(x_src), ptr(b_src, torch.long), ptr(b_q, torch.long)
    d.n_scales = n_scales
    for s in range(n_scales + 1):
        d.src_off[s] = int(src_off[s])
    for s, r in enumerate(radii):
        d.r[s] = -1.0 if r is None else float(r)
    d.max_num_neighbors, d.ns_lo, d.ns_hi = int(max_num_neighbors), float(ns_cut[0]), float(ns_cut[1])
    x_dst = torch.empty(n_dst, 3, dtype=torch.float32, device=dev)
    row_ptr = torch.empty(n_scales * n_dst + 1, dtype=torch.int32, device=dev)
    counts = torch.empty(max(1, n_scales * n_dst), dtype=torch.int32, device=dev)
    n_edges_dev = torch.empty(1, dtype=torch.int32, device=dev)
    d.x_dst, d.row_ptr, d.counts = ptr(x_dst), ptr(row_ptr, torch.int32), ptr(counts, torch.int32)
    d.n_edges, d.cta_sum, d.barrier = ptr(n_edges_dev, torch.int32), ptr(cta_sum, torch.int32), ptr(barrier, torch.int32)
    d.stage_early = 1 if static_sources else 0
    if rows_all is not None:
        d.step, d.n_steps, d.rows_all, d.rows_cur, d.rows_k = ptr(step, torch.int32), rows_all.shape[1], ptr(rows_all), ptr(rows_cur), rows_all.shape[2]

    def run(cap: int):
        e_src = torch.empty(cap, dtype=torch.int32, device=dev)
        e_dst = torch.empty(cap, dtype=torch.int32, device=dev)
        length = torch.empty(cap, dtype=torch.float32, device=dev)
        sh = torch.empty(cap, 9, dtype=torch.float32, device=dev)
        logit = torch.empty(cap, dtype=torch.float32, device=dev)
        d.capacity = cap
        d.edge_src, d.edge_dst = ptr(e_src, torch.int32), ptr(e_dst, torch.int32)
        d.length, d.sh, d.logit = ptr(length), ptr(sh), ptr(logit)
        _call("dedf_head_front", C.byref(d), stream())
        return e_src, e_dst, length, sh, logit

    if capacity is not None:
        cap = max(1, int(capacity))
        d.overflow = ptr(overflow, torch.int32)
        e_src, e_dst, length, sh, logit = run(cap)
        return Csr(row_ptr, e_src, e_dst, n_edges_dev, cap, n_dst, n_scales), length, sh, logit, x_dst
    # eager: guess a capacity, read the true count back (the one host sync of a graph build), re-run if the guess was too small
    ovf = torch.zeros(1, dtype=torch.int32, device=dev)
    d.overflow = ptr(ovf, torch.int32)
    bufs = [run(max(1024, 64 * n_dst))]

    def read_count() -> int:
        # the device clamps the count to the capacity: the overflow flag says whether it is the true one
        n, over = torch.cat([n_edges_dev, ovf]).tolist()
        if over:
            n = int(counts.sum().item())
            ovf.zero_()
            bufs[0] = run(max(1, n))
        return int(n)

    n_edges = plan_value(read_count)
    e_src, e_dst, length, sh, logit = bufs[0]
    E = max(1, n_edges)
    return (Csr(row_ptr, e_src[:n_edges], e_dst[:n_edges], n_edges_dev, n_edges, n_dst, n_scales), length[:E], sh[:E], logit[:E], x_dst)


def radius(x: torch.Tensor, y: torch.Tensor, r: float, batch_x: Optional[torch.Tensor] = None,
           batch_y: Optional[torch.Tensor] = None, max_num_neighbors: int = 32) -> torch.Tensor:
    """Drop-in for torch_cluster.radius: LongTensor[2, E] = (y index, x index), sorted by y then x."""
    g = radius_csr(x, y, [r], b_src=batch_x, b_dst=batch_y, max_num_neighbors=max_num_neighbors)
    return torch.stack([g.edge_dst.long(), g.edge_src.long()], dim=0)


def radius_graph(x: torch.Tensor, r: float, batch: Optional[torch.Tensor] = None, loop: bool = False,
                 max_num_neighbors: int = 32) -> torch.Tensor:
    """Drop-in for torch_cluster.radius_graph (row 0 = centre / destination, row 1 = neighbour / source)."""
    g = radius_csr(x, x, [r], b_src=batch, b_dst=batch, excl_mode=0 if loop else 2,
                   max_num_neighbors=max_num_neighbors if loop else max_num_neighbors + 1)
    return torch.stack([g.edge_dst.long(), g.edge_src.long()], dim=0)


# ----------------------------------------------------------------------------
# per-edge
# ----------------------------------------------------------------------------
def edge_geom(x_src: torch.Tensor, x_dst: torch.Tensor, g: Csr, radii: Optional[Sequence[Optional[float]]] = None,
              src_off: Optional[Sequence[int]] = None, ns_cut: Optional[Tuple[float, float]] = None,
              want_logit: bool = False):
    """-> (length (E), sh (E,9), logit (E) or None)."""
    dev = x_src.device
    E = max(1, g.n_edges)
    length = torch.empty(E, dtype=torch.float32, device=dev)
    sh = torch.empty(E, 9, dtype=torch.float32, device=dev)
    logit = torch.empty(E, dtype=torch.float32, device=dev) if want_logit else None
    n_scales = len(radii) if radii is not None else 1
    so = L.int_array(src_off if src_off is not None else [0, x_src.shape[0]])
    rr = L.float_array([-1.0 if r is None else float(r) for r in (radii if radii is not None else [None])])
    lo, hi = ns_cut if ns_cut is not None else (0.0, -1.0)
    _call("dedf_edge_geom", ptr(x_src.contiguous()), ptr(x_dst.contiguous()), ptr(g.edge_src, torch.int32),
                             ptr(g.edge_dst, torch.int32), ptr(g.n_edges_dev, torch.int32), g.n_edges, n_scales, so, rr,
                             lo, hi, ptr(length), ptr(sh), ptr(logit), stream())
    return length, sh, logit


USE_VALUE_REDUCE = True     # value path reassociated: linear once per destination (dedf_value_reduce) instead of once per edge
USE_TC_TPACT = True     # attention logits + gated values: block-diagonal linear on the tensor cores (dedf_edge_tp_act_tc)
# fp16 hi / lo operand split on kind::f16 instead of the tf32 split (same accuracy for |operands| < 65504, half the shared-memory
# traffic: include/dedf.h, DEDF_TPACT_F16); DEDF_TPACT_F16=0 selects the tf32 split
TPACT_F16 = os.environ.get("DEDF_TPACT_F16", "1") != "0"
MLP_F16 = os.environ.get("DEDF_MLP_F16", "1") != "0"        # the same for dedf_edge_mlp_tc (dedf_mlp_desc.tc_f16)
USE_TC_MLP = True       # per-edge MLPs on the tcgen05 tensor cores (fp16 or tf32 hi / lo split: MLP_F16 below) where the layer widths allow it


def edge_mlp(desc: L.MlpDesc, max_edges: int) -> None:
    _call("dedf_edge_mlp", C.byref(desc), max_edges, stream())


def edge_mlp_tc(desc: L.MlpDesc, max_edges: int) -> None:
    _flops("dedf_edge_mlp_tc", max_edges * 2.0 * sum(desc.dims[i] * desc.dims[i + 1] for i in range(desc.n_layers)))
    _call("dedf_edge_mlp_tc", C.byref(desc), max_edges, stream())


def edge_tp_lin(mul1: int, epilogue: int, x_src: torch.Tensor, x_dst: Optional[torch.Tensor], per_edge_x: bool, g: Csr,
                sh: torch.Tensor, w: torch.Tensor, w_stride: int, W0, W1, W2, bias0, alpha_dot=None, edge_logit=None,
                logits=None, out=None) -> None:
    _call("dedf_edge_tp_lin", mul1, epilogue, ptr(x_src), ptr(x_dst), 1 if per_edge_x else 0,
                                    ptr(g.edge_src, torch.int32), ptr(g.edge_dst, torch.int32),
                                    ptr(g.n_edges_dev, torch.int32), g.n_edges, ptr(sh), ptr(w), w_stride, ptr(W0), ptr(W1),
                                    ptr(W2), ptr(bias0), ptr(alpha_dot), ptr(edge_logit), ptr(logits), ptr(out), stream())


def edge_tp_act_tc(mul1: int, x_src: torch.Tensor, x_dst: Optional[torch.Tensor], g: Csr, sh: torch.Tensor, w: torch.Tensor,
                   w_stride: int, W_tc: torch.Tensor, bias0, alpha_dot, edge_logit, logits: torch.Tensor, out: torch.Tensor,
                   w_perm: bool = False, f16: bool = False) -> None:
    """dedf_edge_tp_lin(EPI_ACT) with the linear layer on the tcgen05 tensor cores.  ``w_perm``: the columns of ``w`` are in
    the kernel's chunk-major order (layers.tp_act_w_perm); ``f16``: ``W_tc`` is the fp16 hi / lo pack (kind::f16 variant)."""
    _flops("dedf_edge_tp_act_tc", g.n_edges * _tp_act_flops(mul1))
    _call("dedf_edge_tp_act_tc", mul1, ptr(x_src), ptr(x_dst), ptr(g.edge_src, torch.int32), ptr(g.edge_dst, torch.int32),
          ptr(g.n_edges_dev, torch.int32), g.n_edges, ptr(sh), ptr(w), w_stride, (1 if w_perm else 0) | (2 if f16 else 0), ptr(W_tc), ptr(bias0), ptr(alpha_dot),
          ptr(edge_logit), ptr(logits), ptr(out), stream())


def segment_softmax_reduce(g: Csr, logits: torch.Tensor, val: torch.Tensor, irr: Tuple[int, int, int]) -> torch.Tensor:
    out = torch.empty(g.n_dst, irr[0] + 3 * irr[1] + 5 * irr[2], dtype=torch.float32, device=val.device)
    _call("dedf_segment_softmax_reduce", ptr(g.row_ptr, torch.int32), g.n_dst, g.n_seg, ptr(logits), ptr(val),
                                               irr[0], irr[1], irr[2], ptr(out), stream())
    return out


K1_PERSIST_BYTES = None        # opt-in (bytes): pin a gathered feature table at least this big in L2 (dedf_l2_persist); measured slower, DESIGN 8


def l2_persist(t: Optional[torch.Tensor]) -> None:
    """Keep ``t`` resident in L2 for the kernels launched into the current stream from now on; ``None`` removes the window."""
    lib = L.load()
    rc = lib.dedf_l2_persist(ptr(t) if t is not None else None, t.numel() * t.element_size() if t is not None else 0, stream())
    L.check(rc, "dedf_l2_persist")


def edge_tp_reduce(mul1: int, x: torch.Tensor, row_ptr: torch.Tensor, edge_src: torch.Tensor, sh: torch.Tensor,
                   w: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    """K1: out[d] = sum_{e->d} alpha[e, head(u)] * DTP(x[src_e], sh_e, w_e)  -> (N_dst, 49 * mul1).
    ``sh`` is (E, 9) (plain-load kernel) or (E, 12) zero-padded rows (TMA bulk-copy pipeline, the fast path)."""
    n_dst = row_ptr.numel() - 1
    assert sh.shape[1] in (9, 12)
    out = torch.empty(n_dst, 49 * mul1, dtype=torch.float32, device=x.device)
    persist = K1_PERSIST_BYTES is not None and x.numel() * 4 >= K1_PERSIST_BYTES and not torch.cuda.is_current_stream_capturing()
    if persist:
        l2_persist(x)
    _call("dedf_edge_tp_reduce", mul1, ptr(x), ptr(row_ptr, torch.int32), ptr(edge_src, torch.int32), ptr(sh), sh.shape[1], ptr(w),
                                       ptr(alpha), n_dst, ptr(out), stream())
    if persist:
        l2_persist(None)
    return out


# ----------------------------------------------------------------------------
# per-node
# ----------------------------------------------------------------------------
def node_linear(x: torch.Tensor, irr_in, irr_out, W: Sequence[Optional[torch.Tensor]], bias0: Optional[torch.Tensor],
                ln: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, ln_eps: float = 1e-5, gate: bool = False,
                res: Optional[torch.Tensor] = None, res_scale: float = 1.0) -> torch.Tensor:
    n = x.shape[0]
    fy = irr_out[0] + 3 * irr_out[1] + 5 * irr_out[2]
    if gate:
        fy -= irr_out[1] + irr_out[2]
    y = torch.empty(n, fy, dtype=torch.float32, device=x.device)
    ln_w, ln_b = (ln if ln is not None else (None, None))
    _flops("dedf_node_linear", n * 2.0 * _irr_mac(irr_in, irr_out))
    _call("dedf_node_linear", ptr(x), n, L.int_array(irr_in), L.int_array(irr_out), ptr(W[0]), ptr(W[1]), ptr(W[2]),
                                    ptr(bias0), ptr(ln_w), ptr(ln_b), ln_eps, 1 if gate else 0, ptr(res), res_scale, ptr(y),
                                    stream())
    return y


USE_NODE_CHAIN = True   # proj -> (+res) -> LN -> fctp_1 -> gate -> fctp_2 -> +res in one launch (dedf_node_chain) instead of three
USE_LINEAR_PAIR = True  # linear_src / linear_dst of a UNet block in one launch (dedf_node_linear_pair)


def node_chain_ok(irr_emb, irr_pre) -> bool:
    """Can dedf_node_chain run these irreps?  (every multiplicity a positive multiple of 4, gate scalars left over)"""
    ms = list(irr_emb) + list(irr_pre) + [irr_pre[0] - irr_pre[1] - irr_pre[2]]
    return USE_NODE_CHAIN and all(m > 0 and m % 4 == 0 for m in ms)


def node_chain(x: torch.Tensor, irr_emb, irr_pre, P, pb, res1: Optional[torch.Tensor], ln_w, ln_b, ln_eps: float,
               A, ab, B, bb) -> torch.Tensor:
    """y1 = proj(x) + pb (+ res1);  y = y1 + fctp_2(Gate(fctp_1(LN(y1))))  -- see include/dedf.h (dedf_node_chain)."""
    d = L.NodeChainDesc()
    y = torch.empty_like(x)
    d.x, d.n = ptr(x), x.shape[0]
    for i in range(3):
        d.irr_emb[i], d.irr_pre[i] = int(irr_emb[i]), int(irr_pre[i])
    d.P0, d.P1, d.P2, d.pb = ptr(P[0]), ptr(P[1]), ptr(P[2]), ptr(pb)
    d.res1 = ptr(res1)
    d.ln_w, d.ln_b, d.ln_eps = ptr(ln_w), ptr(ln_b), float(ln_eps)
    d.A0, d.A1, d.A2, d.ab = ptr(A[0]), ptr(A[1]), ptr(A[2]), ptr(ab)
    d.B0, d.B1, d.B2, d.bb = ptr(B[0]), ptr(B[1]), ptr(B[2]), ptr(bb)
    d.y = ptr(y)
    mid = (irr_pre[0] - irr_pre[1] - irr_pre[2], irr_pre[1], irr_pre[2])
    _flops("dedf_node_chain", x.shape[0] * 2.0 * (_irr_mac(irr_emb, irr_emb) + _irr_mac(irr_emb, irr_pre) + _irr_mac(mid, irr_emb)))
    _call("dedf_node_chain", C.byref(d), stream())
    return y


def node_linear_pair(xa: torch.Tensor, irr_in_a, Wa, ba, xb: torch.Tensor, irr_in_b, Wb, bb, irr_out) -> Tuple[torch.Tensor, torch.Tensor]:
    """(LinearRS_a(xa), LinearRS_b(xb)) with a common output irreps, one launch."""
    fo = irr_out[0] + 3 * irr_out[1] + 5 * irr_out[2]
    ya = torch.empty(xa.shape[0], fo, dtype=torch.float32, device=xa.device)
    yb = torch.empty(xb.shape[0], fo, dtype=torch.float32, device=xb.device)
    arr = lambda W: (L.c_fp * 3)(ptr(W[0]), ptr(W[1]), ptr(W[2]))
    _flops("dedf_node_linear_pair", 2.0 * (xa.shape[0] * _irr_mac(irr_in_a, irr_out) + xb.shape[0] * _irr_mac(irr_in_b, irr_out)))
    _call("dedf_node_linear_pair", ptr(xa), xa.shape[0], L.int_array(irr_in_a), arr(Wa), ptr(ba), ptr(ya),
          ptr(xb), xb.shape[0], L.int_array(irr_in_b), arr(Wb), ptr(bb), ptr(yb), L.int_array(irr_out), stream())
    return ya, yb


def weight_post(x: torch.Tensor, ln_g, ln_b, w, b, use_sigmoid: bool, mult_logit: Optional[torch.Tensor]) -> torch.Tensor:
    y = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    _call("dedf_weight_post", ptr(x), x.shape[0], x.shape[1], ptr(ln_g), ptr(ln_b), ptr(w), ptr(b), 1 if use_sigmoid else 0,
          ptr(mult_logit), ptr(y), stream())
    return y


def gather_rows(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    y = torch.empty(idx.shape[0], x.shape[1], dtype=torch.float32, device=x.device)
    _call("dedf_gather_rows", ptr(x), ptr(idx, torch.long), idx.shape[0], x.shape[1], ptr(y), stream())
    return y


def add_scale(a: torch.Tensor, b: torch.Tensor, s: float) -> torch.Tensor:
    y = torch.empty_like(a)
    _call("dedf_add_scale", ptr(a), ptr(b), s, a.numel(), ptr(y), stream())
    return y


# ----------------------------------------------------------------------------
# score head
# ----------------------------------------------------------------------------
def time_embed(desc: L.TimeDesc, time: torch.Tensor) -> torch.Tensor:
    out = torch.empty(desc.n_scales, time.shape[0], desc.out_dim, dtype=torch.float32, device=time.device)
    _call("dedf_time_embed", C.byref(desc), ptr(time), time.shape[0], ptr(out), stream())
    return out


def query_transform(Ts: torch.Tensor, qx: torch.Tensor, qf: torch.Tensor, irr) -> Tuple[torch.Tensor, torch.Tensor]:
    n_t, n_q = Ts.shape[0], qx.shape[0]
    x_out = torch.empty(n_t * n_q, 3, dtype=torch.float32, device=Ts.device)
    f_out = torch.empty(n_t * n_q, qf.shape[1], dtype=torch.float32, device=Ts.device)
    _call("dedf_query_transform", ptr(Ts), n_t, ptr(qx), ptr(qf), n_q, L.int_array(irr), ptr(x_out), ptr(f_out), stream())
    return x_out, f_out


def score_tp(Ts, qf_rot, key_f, qx, qw, irr, Wd: List[torch.Tensor], Wl0, Wl1, bl, n_vec: int, lin_mult: float):
    n_t = Ts.shape[0]
    ang = torch.empty(n_t, 3, dtype=torch.float32, device=Ts.device)
    lin = torch.empty(n_t, 3, dtype=torch.float32, device=Ts.device)
    arr = lambda ts: (L.c_fp * 2)(ptr(ts[0]), ptr(ts[1]))
    _flops("dedf_score_tp", n_t * qx.shape[0] * SCORE_TP_FLOPS_PER_ROW)
    _call("dedf_score_tp", ptr(Ts), n_t, ptr(qf_rot), ptr(key_f), ptr(qx), ptr(qw), qx.shape[0], L.int_array(irr),
                                 arr(Wd), arr(Wl0), arr(Wl1), arr(bl), n_vec, lin_mult, ptr(ang), ptr(lin), stream())
    return ang, lin


SCORE_TP_FLOPS_PER_ROW = 2.0 * (151_296 + 44_256)      # two 'uvu' tensor products (sparse 3j) + their linear layers, per (pose, query point)


def score_tp_step(Ts, qf, key_f, qx, qw, irr, Wd: List[torch.Tensor], Wl0, Wl1, bl, n_vec: int, lin_mult: float, state=None):
    """dedf_score_tp with the feature rotation D(q) psi applied inside (``qf`` = un-rotated query features) and, with ``state``
    (denoise.StepState), the float64 Langevin update of the poses fused behind it.  -> (ang, lin)."""
    n_t = Ts.shape[0]
    ang = torch.empty(n_t, 3, dtype=torch.float32, device=Ts.device)
    lin = torch.empty(n_t, 3, dtype=torch.float32, device=Ts.device)
    d = L.ScoreStepDesc()
    d.Ts, d.n_t, d.qf, d.key_f, d.qx, d.qw, d.n_q = ptr(Ts), n_t, ptr(qf), ptr(key_f), ptr(qx), ptr(qw), qx.shape[0]
    for i in range(3):
        d.irr[i] = int(irr[i])
    for i in range(2):
        d.Wd[i], d.Wl0[i], d.Wl1[i], d.bl[i] = ptr(Wd[i]), ptr(Wl0[i]), ptr(Wl1[i]), ptr(bl[i])
    d.n_vec, d.lin_mult, d.ang_out, d.lin_out = int(n_vec), float(lin_mult), ptr(ang), ptr(lin)
    if state is not None:
        d.T64, d.sched, d.n_steps = ptr(state.T64, torch.float64), ptr(state.sched, torch.float64), state.sched.shape[0]
        d.counter, d.noise = ptr(state.counter, torch.int32), ptr(state.noise, torch.float64)
        d.seed, d.seed_dev = 0, ptr(state.seed, torch.int64)
        d.ang_mult, d.lin_mult_d = float(state.ang_mult), float(state.lin_mult)
        d.traj, d.T32, d.ticket = ptr(state.traj, torch.float64), ptr(Ts), ptr(state.ticket, torch.int32)
    _flops("dedf_score_tp_step", n_t * qx.shape[0] * SCORE_TP_FLOPS_PER_ROW)
    _call("dedf_score_tp_step", C.byref(d), stream())
    return ang, lin


def pose_update(T: torch.Tensor, ang: torch.Tensor, lin: torch.Tensor, noise: Optional[torch.Tensor], seed: int, offset: int,
                t: float, ang_mult: float, lin_mult: float, alpha_ang: float, alpha_lin: float, temperature: float,
                traj_out: Optional[torch.Tensor], T_f32_out: Optional[torch.Tensor], dev_row: Optional[torch.Tensor] = None,
                dev_counter: Optional[torch.Tensor] = None) -> None:
    _call("dedf_pose_update", ptr(T, torch.float64), T.shape[0], ptr(ang), ptr(lin), ptr(noise, torch.float64), seed, offset,
          t, ang_mult, lin_mult, alpha_ang, alpha_lin, temperature, ptr(traj_out, torch.float64), ptr(T_f32_out),
          ptr(dev_row, torch.float64), ptr(dev_counter, torch.int32), stream())


def sample_advance(sched: torch.Tensor, counter: torch.Tensor, time_out: torch.Tensor, cur_row: torch.Tensor,
                   rows_all: Optional[torch.Tensor] = None, rows_cur: Optional[torch.Tensor] = None) -> None:
    ns, k = (rows_all.shape[0], rows_all.shape[2]) if rows_all is not None else (0, 0)
    _call("dedf_sample_advance", ptr(sched, torch.float64), sched.shape[0], ptr(counter, torch.int32), ptr(time_out),
          ptr(cur_row, torch.float64), ptr(rows_all), ptr(rows_cur), ns, k, stream())


def ebm_energy(key_f: torch.Tensor, query_f: torch.Tensor, qw: torch.Tensor, n_t: int, n_q: int, scale: float) -> torch.Tensor:
    """energy[t] = scale * sum_q w_q |key_f[t,q] - query_f[t,q]|^2   (score_head_ebm.py:171-172)."""
    out = torch.empty(n_t, dtype=torch.float32, device=key_f.device)
    _call("dedf_ebm_energy", ptr(key_f.contiguous()), ptr(query_f.contiguous()), ptr(qw), n_t, n_q, key_f.shape[1], scale, ptr(out), stream())
    return out


def edge_gather_scalar(w: torch.Tensor, g: Csr) -> torch.Tensor:
    """(E, 1) per-edge factor w[edge_src[e]] (0 beyond the true edge count when the buffers are capacity-sized)."""
    E = max(1, g.n_edges)
    out = torch.empty(E, 1, dtype=torch.float32, device=w.device)
    _call("dedf_edge_gather_scalar", ptr(w.contiguous()), ptr(g.edge_src, torch.int32), ptr(g.n_edges_dev, torch.int32), g.n_edges, ptr(out), stream())
    return out


def row_scale(x: torch.Tensor, factor: torch.Tensor, irr) -> torch.Tensor:
    """x[r, :] * factor[r, 0]."""
    y = torch.empty_like(x)
    _call("dedf_group_scale", ptr(x.contiguous()), ptr(factor.contiguous()), x.shape[0], L.int_array(irr), 2, ptr(y), stream())
    return y


def value_reduce(mul1: int, g: Csr, v: torch.Tensor, sh: torch.Tensor, logits: torch.Tensor, post: Optional[torch.Tensor], wv: torch.Tensor,
                 V0: torch.Tensor, V1: torch.Tensor, V2: torch.Tensor, vb: Optional[torch.Tensor]) -> torch.Tensor:
    """out[d] = lin(sum_e softmax(logits)_e,h (x post_e) * dtp(v_e, sh_e, wv)) + bias * sum_e alpha  -> (n_dst, F)."""
    out = torch.empty(g.n_dst, v.shape[1], dtype=torch.float32, device=v.device)
    fe, fd = _value_flops(mul1)
    _flops("dedf_value_reduce", g.n_edges * fe + g.n_dst * fd)
    _call("dedf_value_reduce", mul1, ptr(g.row_ptr, torch.int32), g.n_dst, g.n_seg, ptr(v), ptr(sh), ptr(logits), ptr(post), ptr(wv),
          ptr(V0), ptr(V1), ptr(V2), ptr(vb), ptr(out), stream())
    return out
