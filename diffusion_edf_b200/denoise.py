"""The annealed-Langevin loop of ``ScoreModelBase.sample`` (/root/reference/diffusion_edf/score_model_base.py:146-199) as a
replayed CUDA graph of ONE denoise step.

A step is six launches of libdedf.so and nothing else:

    dedf_head_front      poses -> transformed query points -> multi-scale radius search -> CSR -> edge geometry
                         (+ selects this step's precomputed time rows)
    dedf_edge_mlp_tc     length / time embedding -> pre-linear -> RadialProfile -> per-edge tensor-product weights
    dedf_edge_tp_act_tc  gather -> depthwise CG tensor product -> linear -> attention logits + gated values
    dedf_value_reduce    softmax over incoming edges, value tensor product, segment reduce, value linear
    dedf_node_chain      proj -> LN -> FFN -> residual
    dedf_score_tp_step   D(q) psi, the two score tensor products, gate, mean, rotate back, orbital term, weighted sum,
                         and the float64 pose update + step counter

Everything a step reads that changes from step to step (poses, step index, schedule row, time rows, noise window) lives in device
memory, so the captured graph has no host-side parameters and is replayed ``sum(N_steps)`` times back to back.

The graph and its static buffers are cached on the model (keyed by the shapes): a second ``sample`` call with the same shapes -- a
server handling the next request -- copies its inputs into the static buffers and replays; nothing is re-captured.

Edge buffers are sized from the edge count of the seeds (x3, at least 96 edges per query node) instead of the worst case; the
kernels clamp to the capacity and raise a device flag, which is read ONCE after the loop: on overflow the capacity is doubled, the
step re-captured and the (deterministic: Philox noise is keyed by (seed, pose, step)) loop run again.
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional, Sequence

import torch

from . import ops
from .gnn_data import FeaturedPoints


class StepState(NamedTuple):
    """Device-resident state of a replayed denoise loop (see include/dedf.h, dedf_score_tp_step / dedf_head_front)."""
    T64: torch.Tensor        # (nT, 7) float64 poses, updated in place
    sched: torch.Tensor      # (n_steps, 4) float64 [t, alpha_ang, alpha_lin, temperature]
    counter: torch.Tensor    # int32 (1) step index
    noise: Optional[torch.Tensor]   # (n_steps, nT, 6) float64 or None (Philox)
    seed: torch.Tensor       # int64 (1) Philox seed (device: a new seed needs no re-capture)
    traj: torch.Tensor       # (n_steps + 2, nT, 7) float64
    ticket: torch.Tensor     # int32 (1) zeroed once
    rows_all: torch.Tensor   # (n_scales, n_steps, K) time rows of the whole schedule
    rows_cur: torch.Tensor   # (n_scales, 1, K) this step's rows
    ang_mult: float
    lin_mult: float


class DenoiseGraph:
    MARGIN, MIN_PER_NODE, PAD = 3.0, 96, 1024
    STEPS_PER_GRAPH = 25      # steps unrolled into one graph launch: a launch costs ~30 us of fixed overhead on top of its
                              # kernels (profiles/r2_step_ablation_*.json), and consecutive steps inside one graph stay linked
                              # by programmatic dependent launch (the front kernel's prologue overlaps the previous step's tail)

    def __init__(self, model, n_t: int, n_steps: int, sources: Sequence, query: FeaturedPoints, with_noise: bool, dev: torch.device):
        self.model, self.n_t, self.n_steps, self.dev = model, n_t, n_steps, dev
        head = model.score_head
        field = head.key_tensor_field
        f64, f32 = torch.float64, torch.float32
        self.T64 = torch.zeros(n_t, 7, dtype=f64, device=dev)
        self.T32 = torch.zeros(n_t, 7, dtype=f32, device=dev)
        self.traj = torch.zeros(n_steps + 2, n_t, 7, dtype=f64, device=dev)
        self.sched = torch.zeros(n_steps, 4, dtype=f64, device=dev)
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        self.seed = torch.zeros(1, dtype=torch.int64, device=dev)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        self.noise = torch.zeros(n_steps, n_t, 6, dtype=f64, device=dev) if with_noise else None
        K = field.fc_neurons[0]
        self.rows_all = torch.zeros(field.n_scales, n_steps, K, dtype=f32, device=dev)
        self.rows_cur = torch.zeros(field.n_scales, 1, K, dtype=f32, device=dev)
        # static copies of the scene field / query points (the graph holds their addresses)
        self.src_off = list(sources[2])
        self.src = [t.detach().clone() if isinstance(t, torch.Tensor) else t for t in sources]
        self.query = FeaturedPoints(x=query.x.detach().clone(), f=query.f.detach().clone(), b=query.b.detach().clone(), w=query.w.detach().clone())
        self.capacity = 0
        self.graph: Optional[torch.cuda.CUDAGraph] = None        # one step
        self.graph_multi: Optional[torch.cuda.CUDAGraph] = None  # STEPS_PER_GRAPH steps
        self.n_multi = min(self.STEPS_PER_GRAPH, n_steps)
        self.n_kernels = 0
        self.replans = 0

    # ------------------------------------------------------------------ one step
    def _state(self) -> StepState:
        return StepState(self.T64, self.sched, self.counter, self.noise, self.seed, self.traj, self.ticket, self.rows_all, self.rows_cur,
                         float(self.model.ang_mult), float(self.model.lin_mult))

    def _step(self) -> None:
        self.model.score_head.denoise_step(self.T32, self.query, self.src, self.capacity, self.overflow, self._state())

    def _reset(self, T_seed: torch.Tensor) -> None:
        self.T64.copy_(T_seed)
        self.T32.copy_(self.T64)
        self.traj[0].copy_(self.T64)
        self.counter.zero_()
        self.overflow.zero_()

    def _capture(self, T_seed: torch.Tensor) -> None:
        """(Re-)capture the step for the current capacity.  One eager step is the warm-up (allocator pools, lazy kernel attributes)."""
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self._step()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self.graph = torch.cuda.CUDAGraph()
        k0 = ops.LAUNCHES
        with torch.cuda.graph(self.graph):
            self._step()
        self.n_kernels = ops.LAUNCHES - k0
        self.graph_multi = None
        if self.n_multi > 1:
            self.graph_multi = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_multi):
                for _ in range(self.n_multi):
                    self._step()
        ops.LAUNCHES = k0 + self.n_kernels
        self._reset(T_seed)

    def _initial_capacity(self) -> int:
        """Edge count of the seeds (one eager front launch; the only sizing read of a fresh plan)."""
        head = self.model.score_head
        field = head.key_tensor_field
        ns = field.r_mincut_nonscalar_sh
        g, *_ = ops.head_front(self.T32, self.query.x, self.query.b, self.src[0], self.src[1], self.src_off, field.r_cluster_multiscale,
                               (0.2 * ns, 1.0 * ns))
        n_dst = self.n_t * self.query.x.shape[0]
        worst = n_dst * sum(self.src_off[s + 1] - self.src_off[s] if r is None else min(self.src_off[s + 1] - self.src_off[s], 1000)
                            for s, r in enumerate(field.r_cluster_multiscale))
        return max(1, min(worst, max(int(self.MARGIN * g.n_edges), self.MIN_PER_NODE * n_dst) + self.PAD))

    # ------------------------------------------------------------------ run
    def run(self, T_seed: torch.Tensor, sources: Sequence, query: FeaturedPoints, rows: List[List[float]], rows_all: torch.Tensor,
            noise: Optional[torch.Tensor], seed: int) -> torch.Tensor:
        for dst, src in zip(self.src, sources):
            if isinstance(dst, torch.Tensor):
                dst.copy_(src)
        self.query.x.copy_(query.x); self.query.f.copy_(query.f); self.query.b.copy_(query.b); self.query.w.copy_(query.w)
        self.sched.copy_(torch.tensor(rows, dtype=torch.float64))
        self.rows_all.copy_(rows_all)
        self.seed.fill_(int(seed) & 0x7fffffffffffffff)
        if self.noise is not None:
            self.noise.copy_(noise)
        self._reset(T_seed)
        if self.graph is None:
            self.capacity = self._initial_capacity()
            self._capture(T_seed)
        worst = None
        while True:
            n_full = self.n_steps // self.n_multi if self.graph_multi is not None else 0
            for _ in range(n_full):
                self.graph_multi.replay()
            for _ in range(self.n_steps - n_full * self.n_multi):
                self.graph.replay()
            ops.LAUNCHES += self.n_kernels * self.n_steps
            self.traj[self.n_steps + 1].copy_(self.T64)
            if int(self.overflow.item()) == 0:             # the one host read of a whole denoise loop
                return self.traj.clone()
            # an edge list outgrew the buffers at some step: grow, re-capture, run the (deterministic) loop again
            self.replans += 1
            self.capacity *= 2
            self._reset(T_seed)
            self._capture(T_seed)
