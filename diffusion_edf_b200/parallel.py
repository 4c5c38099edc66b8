"""Multi-GPU plumbing: the pose batch is embarrassingly parallel given the encoded scene field
(score_head.py:153-209 is row-wise in nT; score_model_base.py:178-193 updates rows independently), so the
only data-path collectives are ONE broadcast of the packed scene field + query points at set-up and ONE
all-gather of the resulting rows at tear-down.  Zero collectives per diffusion step.
One process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .gnn_data import FeaturedPoints


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous chunk [lo, hi) of ``n`` rows owned by ``rank`` (chunks of ceil(n / world))."""
    chunk = (n + world - 1) // world
    lo = min(rank * chunk, n)
    return lo, min(lo + chunk, n)


def _pack(keys: Sequence[FeaturedPoints], query: FeaturedPoints) -> Tuple[torch.Tensor, torch.Tensor]:
    F = keys[0].f.shape[1]
    header = torch.tensor([len(keys), F, query.x.shape[0]] + [k.x.shape[0] for k in keys], dtype=torch.int64)
    rows = [torch.cat([k.x, k.f, k.b.to(k.x.dtype).unsqueeze(1)], dim=1) for k in keys]
    qw = query.w if query.w is not None else torch.ones(query.x.shape[0], dtype=query.x.dtype, device=query.x.device)
    rows.append(torch.cat([query.x, query.f, qw.unsqueeze(1)], dim=1))
    return header, torch.cat(rows, dim=0).contiguous()


def _unpack(header: torch.Tensor, payload: torch.Tensor) -> Tuple[List[FeaturedPoints], FeaturedPoints]:
    h = header.tolist()
    n_scales, F, n_q = h[0], h[1], h[2]
    keys, off = [], 0
    for n in h[3:3 + n_scales]:
        blk = payload[off:off + n]
        keys.append(FeaturedPoints(x=blk[:, :3].contiguous(), f=blk[:, 3:3 + F].contiguous(), b=blk[:, 3 + F].to(torch.long), w=None))
        off += n
    blk = payload[off:off + n_q]
    query = FeaturedPoints(x=blk[:, :3].contiguous(), f=blk[:, 3:3 + F].contiguous(),
                           b=torch.zeros(n_q, dtype=torch.long, device=payload.device), w=blk[:, 3 + F].contiguous())
    return keys, query


def broadcast_scene_field(keys: Optional[Sequence[FeaturedPoints]], query: Optional[FeaturedPoints], src: int = 0,
                          device: Optional[torch.device] = None, sizes: Optional[Sequence[int]] = None,
                          ) -> Tuple[List[FeaturedPoints], FeaturedPoints]:
    """Rank ``src`` holds the encoded multiscale key field and the query points; every rank returns them.

    The payload [sum_s N_s + nQ, 3 + F + 1] fp32 (coordinates | features | batch id or query weight) goes out in a
    single broadcast (~2.45 MB for a 10 k-point scene).  ``sizes`` = [n_scales, F, nQ, N_0, ...] lets the receivers
    skip the small header broadcast when the shapes are known up front."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return list(keys), query
    rank = dist.get_rank()
    if rank == src:
        header, payload = _pack(keys, query)
        device = payload.device if device is None else device
        payload = payload.to(device)
    assert device is not None
    if sizes is None:
        hdr = (header if rank == src else torch.zeros(3 + 8, dtype=torch.int64))
        if rank == src:
            hdr = torch.cat([header, torch.zeros(11 - len(header), dtype=torch.int64)])
        hdr = hdr.to(device)
        dist.broadcast(hdr, src=src)
        hdr = hdr.cpu()
    else:
        hdr = torch.tensor(list(sizes), dtype=torch.int64)
    if rank != src:
        n_rows = int(hdr[3:3 + int(hdr[0])].sum() + hdr[2])
        payload = torch.empty(n_rows, 3 + int(hdr[1]) + 1, dtype=torch.float32, device=device)
    dist.broadcast(payload, src=src)
    return _unpack(hdr, payload)


def all_gather_rows(mine: torch.Tensor, n_total: int) -> torch.Tensor:
    """Inverse of ``shard_range``: concatenates every rank's rows (chunks padded to equal size for the collective)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return mine
    world = dist.get_world_size()
    chunk = (n_total + world - 1) // world
    pad = torch.zeros((chunk,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
    pad[:mine.shape[0]] = mine
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat(out, dim=0)[:n_total]


def sharded_forward(model, Ts_local: torch.Tensor, time_local: torch.Tensor, key_pcd: Optional[FeaturedPoints],
                    query_pcd: FeaturedPoints, src: int = 0, sizes: Optional[Sequence[int]] = None):
    """``MultiscaleScoreModel.forward`` with the pose batch sharded over the ranks: rank ``src`` encodes the scene
    (UNet) and the query points once, the packed field is broadcast, every rank scores its own poses.
    Returns this rank's (ang (nT_local,3), lin (nT_local,3))."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        (ang, lin), _ = model(Ts_local, time_local, key_pcd, query_pcd)
        return ang, lin
    if rank == src:
        keys = model.get_key_pcd_multiscale(key_pcd)
        query = model.get_query_pcd(query_pcd)
        query = FeaturedPoints(query.x.detach(), query.f.detach(), query.b, query.w.detach())
    else:
        keys, query = None, None
    keys, query = broadcast_scene_field(keys, query, src=src, device=Ts_local.device, sizes=sizes)
    return model.score_head(Ts=Ts_local, key_pcd_multiscale=keys, query_pcd=query, time=time_local)


def sharded_sample(model, T_seed: torch.Tensor, key_pcd: Optional[FeaturedPoints], query_pcd: Optional[FeaturedPoints],
                   src: int = 0, gather: bool = True, **sample_kwargs) -> torch.Tensor:
    """``ScoreModelBase.sample`` with the seeds sharded over the ranks (BASELINE config C3): rank ``src`` encodes the scene
    and the query points, ONE broadcast ships the packed field, every rank denoises its contiguous chunk of ``T_seed`` with
    its own Philox stream (zero collectives per diffusion step), ONE all-gather returns the trajectories (steps+2, nT, 7)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if rank == src:
        keys = model.get_key_pcd_multiscale(key_pcd)
        query = model.get_query_pcd(query_pcd)
        query = FeaturedPoints(query.x.detach(), query.f.detach(), query.b, query.w.detach())
    else:
        keys, query = None, None
    if world > 1:
        keys, query = broadcast_scene_field(keys, query, src=src, device=T_seed.device)
    lo, hi = shard_range(T_seed.shape[0], rank, world)
    base_seed = int(getattr(model, "sample_seed", 0))
    model.sample_seed = base_seed * 1000003 + rank      # independent noise per rank (restored below: calls do not compound)
    try:
        traj = model.sample(T_seed[lo:hi].contiguous(), keys, query, **sample_kwargs)    # (S, n_local, 7)
    finally:
        model.sample_seed = base_seed
    if world == 1 or not gather:
        return traj
    rows = all_gather_rows(traj.transpose(0, 1).contiguous(), T_seed.shape[0])       # (nT, S, 7)
    return rows.transpose(0, 1).contiguous()


def allreduce_gradients(params, average: bool = True, bucket_bytes: int = 32 << 20) -> int:
    """Data-parallel training step (BASELINE config C5: one synthetic demo per rank): sum (or average) the gradients of
    ``params`` over all ranks.  Gradients are flattened into buckets of ~``bucket_bytes`` so that the ~1.8 M fp32
    parameters (7.4 MB) travel in ONE all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests).  Parameters whose
    gradient is None on this rank contribute zeros (every rank must issue the same collectives).  Returns the number of
    collectives issued."""
    params = [p for p in params if p.requires_grad]
    if not dist.is_initialized() or dist.get_world_size() == 1 or not params:
        return 0
    world = dist.get_world_size()
    n_coll, bucket, size = 0, [], 0

    def flush():
        nonlocal n_coll, bucket, size
        if not bucket:
            return
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat.div_(world)
        off = 0
        for p in bucket:
            n = p.numel()
            g = flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
        n_coll += 1
        bucket, size = [], 0

    for p in params:
        bucket.append(p)
        size += p.numel() * p.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    return n_coll
