"""ScoreModelBase on the CUDA path: ``forward`` / ``sample`` / ``get_train_loss`` with the reference's
signatures (/root/reference/diffusion_edf/score_model_base.py:22-225)."""
from __future__ import annotations

import os

import math
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
from torch import nn

from . import ops
from .gnn_data import FeaturedPoints, detach_featured_points
from .graphs import GraphedCallable


class ScoreModelBase(nn.Module):
    lin_mult: float
    ang_mult: float

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.register_buffer("q_indices", torch.tensor([[1, 2, 3], [0, 3, 2], [3, 0, 1], [2, 1, 0]], dtype=torch.long), persistent=False)
        self.register_buffer("q_factor", torch.tensor([[-0.5, -0.5, -0.5], [0.5, -0.5, 0.5], [0.5, 0.5, -0.5], [-0.5, 0.5, 0.5]]), persistent=False)
        self.sample_seed = 0
        # CUDA-graph replay of forward() / the denoise step when no gradients are needed (see graphs.py)
        self.use_cuda_graph = True
        self._graphs = {}
        self._denoise_graphs = {}
        self._prefetch_tab = None

    # ------------------------------------------------------------------ reduced-precision switch of the reference
    def half(self):
        """agent.py:48-51 calls ``model.half()`` when ``half_precision`` is configured.  The kernels compute in fp32 (never
        below the reference's fp16 arithmetic), so the parameters stay fp32: the switch only changes the interface --
        fp16 inputs (poses, times, point clouds) are widened on entry and the scores / features come back in fp16, as a
        caller of the reference's half model sees them."""
        self._io_half = True
        return self

    def float(self):
        self._io_half = False
        return super().float()

    @staticmethod
    def _widen(v):
        """fp16 / bf16 tensor or FeaturedPoints (or a list of them) -> fp32; anything else unchanged."""
        if isinstance(v, torch.Tensor):
            return v.float() if v.dtype in (torch.float16, torch.bfloat16) else v
        if isinstance(v, FeaturedPoints):
            return FeaturedPoints(x=ScoreModelBase._widen(v.x), f=ScoreModelBase._widen(v.f), b=v.b, w=ScoreModelBase._widen(v.w))
        if isinstance(v, (list, tuple)):
            return type(v)(ScoreModelBase._widen(e) for e in v)
        return v

    def _narrow(self, v):
        """fp32 result -> fp16 when the model was switched with .half()."""
        if not getattr(self, "_io_half", False):
            return v
        if isinstance(v, torch.Tensor):
            return v.half() if v.dtype == torch.float32 else v
        if isinstance(v, FeaturedPoints):
            return FeaturedPoints(x=self._narrow(v.x), f=self._narrow(v.f), b=v.b, w=self._narrow(v.w))
        if isinstance(v, (list, tuple)):
            return type(v)(self._narrow(e) for e in v)
        return v

    # subclasses implement the fp32 internals; the public methods add the .half() interface conversion
    def _key_pcd_multiscale(self, pcd: FeaturedPoints) -> List[FeaturedPoints]:
        raise NotImplementedError

    def _query_pcd(self, pcd: FeaturedPoints) -> FeaturedPoints:
        raise NotImplementedError

    def get_key_pcd_multiscale(self, pcd: FeaturedPoints) -> List[FeaturedPoints]:
        return self._narrow(self._key_pcd_multiscale(self._widen(pcd)))

    def get_query_pcd(self, pcd: FeaturedPoints) -> FeaturedPoints:
        return self._narrow(self._query_pcd(self._widen(pcd)))

    # ------------------------------------------------------------------ training loss (forward value)
    def get_train_loss(self, Ts, time, key_pcd, query_pcd, target_ang_score, target_lin_score):
        """The reference's loss, statistics and (when parameters require grad) a differentiable graph
        (score_model_base.py:41-107).  With gradients enabled the model runs through the training path
        (train_path.py: un-fused CUDA primitives with hand-written backward kernels); otherwise through the fused
        inference kernels."""
        assert target_ang_score.ndim == 2 and target_ang_score.shape[-1] == 3
        assert target_lin_score.ndim == 2 and target_lin_score.shape[-1] == 3
        assert len(time) == len(target_ang_score) == len(target_lin_score)
        Ts, time, key_pcd, query_pcd, target_ang_score, target_lin_score = self._widen((Ts, time, key_pcd, query_pcd, target_ang_score,
                                                                                        target_lin_score))
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad:
            from . import train_path
            from .keypoint_extractor import KeypointExtractor
            if isinstance(self.key_model, KeypointExtractor):                        # PointAttentiveScoreModel
                key_ms = [train_path.keypoint_extractor(self.key_model, key_pcd)]
            else:
                key_ms = train_path.unet_forward(self.key_model, key_pcd)
            if isinstance(self.query_model, KeypointExtractor):
                q = train_path.keypoint_extractor(self.query_model, query_pcd)      # place configs
            else:
                q = self._query_pcd(query_pcd)                                       # StaticKeypointModel: torch views of parameters
            ang, lin = train_path.score_head(self.score_head, Ts, key_ms, q, time)
        else:
            key_ms = self._key_pcd_multiscale(key_pcd)
            q = self._query_pcd(query_pcd)
            ang, lin = self.score_head(Ts=Ts, key_pcd_multiscale=key_ms, query_pcd=q, time=time)
        t_ang = target_ang_score * torch.sqrt(time[..., None]) * self.ang_mult
        t_lin = target_lin_score * torch.sqrt(time[..., None]) * self.lin_mult
        ang_loss = torch.sum(torch.square(t_ang - ang), dim=-1).mean(dim=-1)
        lin_loss = torch.sum(torch.square(t_lin - lin), dim=-1).mean(dim=-1)
        loss = ang_loss + lin_loss
        tn_a, tn_l = torch.norm(t_ang, dim=-1), torch.norm(t_lin, dim=-1)
        sn_a, sn_l = torch.norm(ang, dim=-1), torch.norm(lin, dim=-1)
        dp_a, dp_l = (ang * t_ang).sum(-1), (lin * t_lin).sum(-1)
        stats = torch.stack([loss, ang_loss, lin_loss, tn_a.mean(), tn_l.mean(), sn_a.mean(), sn_l.mean(), dp_a.mean(),
                             dp_l.mean(), (dp_a / tn_a / sn_a).mean(), (dp_l / tn_l / sn_l).mean()]).tolist()   # one D2H
        names = ["Loss/train", "Loss/angular", "Loss/linear", "norm/target_ang", "norm/target_lin", "norm/inferred_ang",
                 "norm/inferred_lin", "alignment/unnormalized/ang", "alignment/unnormalized/lin",
                 "alignment/normalized/ang", "alignment/normalized/lin"]
        statistics = dict(zip(names, stats))
        fp_info = {"key_fp": None, "query_fp": detach_featured_points(q)}
        tensor_info = {"ang_score": ang.detach(), "lin_score": lin.detach()}
        return loss, fp_info, tensor_info, statistics

    # ------------------------------------------------------------------ sampling
    @torch.no_grad()
    def sample(self, T_seed: torch.Tensor, scene_pcd_multiscale: List[FeaturedPoints], grasp_pcd: FeaturedPoints,
               diffusion_schedules: List[Union[List[float], Tuple[float, float]]], N_steps: List[int], timesteps: List[float],
               temperatures: Union[Union[int, float], Sequence[Union[int, float]]] = 1.0, log_t_schedule: bool = True,
               time_exponent_temp: float = 0.5, time_exponent_alpha: float = 0.5,
               noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Annealed Langevin dynamics on SE(3) (score_model_base.py:110-204): returns (sum(N_steps)+2, nT, 7) float64.

        The pose integrator runs on the device in float64 (dedf_pose_update); the noise comes from Philox
        (seed = ``self.sample_seed``) unless ``noise`` (sum(N_steps), nT, 6) standard normals is given."""
        if isinstance(temperatures, (int, float)):
            temperatures = [float(temperatures)] * len(diffusion_schedules)
        scene_pcd_multiscale, grasp_pcd = self._widen(list(scene_pcd_multiscale)), self._widen(grasp_pcd)
        dev = T_seed.device
        nT = T_seed.shape[0]
        T = T_seed.detach().to(torch.float64).contiguous().clone()
        total = int(sum(N_steps))
        traj = torch.empty(total + 2, nT, 7, dtype=torch.float64, device=dev)
        traj[0].copy_(T)
        T32 = T.to(torch.float32)
        sources = self.score_head.key_tensor_field.encode_sources(scene_pcd_multiscale)   # pose-independent, once per scene
        # schedule rows [t, alpha_ang, alpha_lin, temperature] in float64 (score_model_base.py:146-171)
        rows = []
        for n, sch in enumerate(diffusion_schedules):
            s0, s1 = float(sch[0]), float(sch[1])
            if log_t_schedule:
                ts = torch.logspace(math.log(s0), math.log(s1), N_steps[n], base=math.e, dtype=torch.float64).tolist()
            else:
                ts = torch.linspace(s0, s1, N_steps[n], dtype=torch.float64).tolist()
            for t in ts:
                rows.append([t, (self.ang_mult ** 2) * (t ** time_exponent_alpha) * timesteps[n],
                             (self.lin_mult ** 2) * (t ** time_exponent_alpha) * timesteps[n],
                             float(temperatures[n]) * (t ** time_exponent_temp)])
        if noise is not None:
            noise = noise.to(torch.float64).contiguous()
        # all poses share the time of a step and the schedule is known up front: ONE time-embedding launch for the whole loop
        rows_all = self.score_head.time_rows_for(torch.tensor([r[0] for r in rows], dtype=torch.float32, device=dev))
        graph_ok = self.use_cuda_graph and ops.USE_HEAD_FRONT and self.score_head._front_ok(scene_pcd_multiscale, sources)
        if not graph_ok:
            for step, (t, a_ang, a_lin, temperature) in enumerate(rows):
                time = torch.full((1,), t, dtype=torch.float32, device=dev)
                ang, lin = self.score_head(Ts=T32, key_pcd_multiscale=scene_pcd_multiscale, query_pcd=grasp_pcd, time=time,
                                           sources=sources, shared_time=True, time_rows=rows_all[:, step:step + 1].contiguous())
                nz = noise[step] if noise is not None else None
                ops.pose_update(T, ang, lin, nz, int(self.sample_seed), step, t, self.ang_mult, self.lin_mult, a_ang, a_lin,
                                temperature, traj[step + 1], T32)
            traj[total + 1].copy_(T)
            return traj
        # ---- one CUDA graph per denoise step, cached across calls (denoise.py): schedule, step counter, poses, trajectory, noise
        # and seed all live on the device; edge buffers are sized from the seeds' edge count with a device-side overflow flag that
        # is read once after the loop.
        from .denoise import DenoiseGraph
        key = (nT, total, tuple(int(v) for v in sources[2]), tuple(grasp_pcd.x.shape), noise is not None, str(dev), len(sources),
               self._param_signature(), ops.USE_TC_MLP, ops.MLP_F16, ops.USE_TC_TPACT, ops.TPACT_F16, ops.USE_VALUE_REDUCE, ops.USE_NODE_CHAIN)
        dg = self._denoise_graphs.get(key)
        if dg is None:
            if len(self._denoise_graphs) >= 2:
                self._denoise_graphs.clear()
            dg = DenoiseGraph(self, nT, total, sources, grasp_pcd, noise is not None, dev)
            self._denoise_graphs[key] = dg
        return dg.run(T, sources, grasp_pcd, rows, rows_all, noise, int(self.sample_seed))

    # ------------------------------------------------------------------ forward
    def _apply(self, fn, *args, **kwargs):
        self.__dict__["_param_cache"] = None               # .to() / .cuda() / .float(): storage moves
        return super()._apply(fn, *args, **kwargs)

    def _param_signature(self):
        """Changes whenever a parameter is updated in place (optimizer step, load_state_dict: version counters only grow) or the
        parameters move (first / last storage pointer).  Cheap on purpose: it is evaluated on EVERY forward to pick the cached CUDA
        graph -- walking the module tree (733 tensors) cost 1.3 ms per call, a quarter of the C2 step (profiles/r2_s8_timeline_*)."""
        ps = self.__dict__.get("_param_cache")
        if ps is None:
            ps = tuple(self.parameters())
            self.__dict__["_param_cache"] = ps
        v = 0
        for p in ps:
            v += p._version
        return (len(ps), ps[0].data_ptr(), ps[-1].data_ptr(), v)

    def _weight_tensors(self):
        """Every tensor the kernels read as a weight: parameters plus the cached kernel-layout copies."""
        out = [p.data for p in self.parameters()]

        def walk(v):
            if isinstance(v, torch.Tensor):
                out.append(v)
            elif isinstance(v, dict):
                for x in v.values():
                    walk(x)
            elif isinstance(v, (list, tuple)):
                for x in v:
                    walk(x)

        for m in self.modules():
            c = getattr(m, "_packed", None)
            if c is not None and getattr(c, "_val", None) is not None:
                walk(c._val)
            for name in ("_pre_tc", "_pre_tc16"):
                if isinstance(getattr(m, name, None), torch.Tensor):
                    out.append(getattr(m, name))
            for name in ("_time_cache", "_tp_cache", "_pre_cache"):      # (key, value) caches of the head / field
                c = getattr(m, name, None)
                if isinstance(c, tuple) and len(c) == 2 and c[1] is not None:
                    walk(c[1])
        return out

    def prefetch_weights(self) -> None:
        """Issue L2 prefetches for all weights (7-15 MB << the 126 MB L2): the forward's few-CTA kernels then hit L2 instead
        of paying DRAM latency per weight row.  The table is rebuilt when parameters or packed copies change; it is built
        outside CUDA-graph capture (the warm-up pass of graphs.py) and only looked up during capture."""
        ts = self._weight_tensors()
        key = tuple(t.data_ptr() for t in ts)
        if self._prefetch_tab is None or self._prefetch_tab[0] != key:
            if torch.cuda.is_current_stream_capturing():
                return
            self._prefetch_tab = (key, ops.prefetch_table(ts))
        ops.prefetch_l2(self._prefetch_tab[1])

    def _forward_tensors(self, Ts, time, kx, kf, kb, qx, qf, qb):
        ops.stamp("forward start")
        self.prefetch_weights()
        # The query model and the time embedding depend on neither the scene nor the key encoder: they run on their own stream
        # under the encoder's farthest-point sampling (the main stream is idle for it) instead of after the encoder (~70 us of
        # launches on the critical path: in-graph timeline, profiles/r2_s11_*).
        early = Ts.is_cuda and os.environ.get("DEDF_EARLY_QUERY", "1") != "0" and hasattr(self.score_head, "time_rows_for")
        q = time_rows = ev = None
        if early:
            main = torch.cuda.current_stream()
            if self.__dict__.get("_aux_stream") is None or self._aux_stream.device != Ts.device:
                self.__dict__["_aux_stream"] = torch.cuda.Stream(device=Ts.device)
            aux = self._aux_stream
            aux.wait_stream(main)
            with torch.cuda.stream(aux):
                q = self._query_pcd(FeaturedPoints(qx, qf, qb))
                time_rows = self.score_head.time_rows_for(time)
                ev = torch.cuda.Event()
                ev.record(aux)
        key_ms = self._key_pcd_multiscale(FeaturedPoints(kx, kf, kb))
        ops.stamp("key encoder done")
        if early:
            main.wait_event(ev)
            for t in (q.x, q.f, q.b, q.w, time_rows):
                if isinstance(t, torch.Tensor):
                    t.record_stream(main)
            out = self.score_head(Ts=Ts, key_pcd_multiscale=key_ms, query_pcd=q, time=time, time_rows=time_rows)
        else:
            q = self._query_pcd(FeaturedPoints(qx, qf, qb))
            out = self.score_head(Ts=Ts, key_pcd_multiscale=key_ms, query_pcd=q, time=time)
        ops.stamp("score head done")
        return out

    def forward(self, Ts: torch.Tensor, time: torch.Tensor, key_pcd: FeaturedPoints, query_pcd: FeaturedPoints,
                debug: bool = False):
        Ts, time, key_pcd, query_pcd = self._widen((Ts, time, key_pcd, query_pcd))
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        graph_ok = getattr(self.query_model, "graph_safe", False)      # see graphs.py: shapes must not depend on the data
        if self.use_cuda_graph and graph_ok and Ts.is_cuda and not debug and not needs_grad and not self.training:
            ins = [Ts.contiguous(), time.contiguous(), key_pcd.x.contiguous(), key_pcd.f.contiguous(), key_pcd.b.contiguous(),
                   query_pcd.x.contiguous(), query_pcd.f.contiguous(), query_pcd.b.contiguous()]
            key = (tuple(tuple(t.shape) for t in ins), str(Ts.device), self._param_signature())
            g = self._graphs.get(key)
            if g is None:
                if len(self._graphs) >= 4:
                    self._graphs.clear()
                g = GraphedCallable(self._forward_tensors, ins)
                self._graphs[key] = g
                return self._narrow(tuple(o.clone() for o in g.eager_out)), None
            return self._narrow(tuple(g(*ins))), None
        key_ms = self._key_pcd_multiscale(key_pcd)
        q = self._query_pcd(query_pcd)
        score = self.score_head(Ts=Ts, key_pcd_multiscale=key_ms, query_pcd=q, time=time)
        dbg = ([detach_featured_points(k) for k in key_ms], detach_featured_points(q)) if debug else None
        return self._narrow(tuple(score)), self._narrow(dbg)
