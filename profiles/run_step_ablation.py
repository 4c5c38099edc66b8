"""Marginal in-graph cost of every kernel of ONE replayed denoise step (denoise.py): the step is captured six times, cut after
its k-th launch (the later entry points are skipped), and each prefix graph is replayed back to back; the differences between
consecutive prefixes are what each kernel adds to the replayed step, PDL overlap included (ncu's serialised cold-cache durations
cannot show that).
    python profiles/run_step_ablation.py [n_poses=128] [replays=300]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusion_edf_b200 import FeaturedPoints, MultiscaleScoreModel, ops
from diffusion_edf_b200.denoise import DenoiseGraph
from diffusion_edf_b200.synthetic import make_poses, make_scene, model_kwargs

n_poses = int(sys.argv[1]) if len(sys.argv) > 1 else 128
replays = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MultiscaleScoreModel(**model_kwargs(), deterministic=True).eval().to(dev)
model.requires_grad_(False)
x, rgb = make_scene(10_000, seed=0)
T_seed, _ = make_poses(n_poses, x, seed=0)
key = FeaturedPoints(x.to(dev), rgb.to(dev), torch.zeros(len(x), dtype=torch.long, device=dev))
grasp = FeaturedPoints(torch.zeros(8, 3, device=dev), torch.zeros(8, 3, device=dev), torch.zeros(8, dtype=torch.long, device=dev))
ORDER = ["dedf_head_front", "dedf_edge_mlp_tc", "dedf_edge_tp_act_tc", "dedf_value_reduce", "dedf_node_chain", "dedf_score_tp_step"]
real_call = ops._call
out = {"n_poses": n_poses, "replays": replays}
with torch.no_grad():
    keys = model.get_key_pcd_multiscale(key)
    q = model.get_query_pcd(grasp)
    src = model.score_head.key_tensor_field.encode_sources(keys)
    n_steps = replays + 64
    rows = [[0.5, 1e-3, 1e-3, 1.0]] * n_steps
    rows_all = model.score_head.time_rows_for(torch.full((n_steps,), 0.5, device=dev))
    prev = 0.0
    for k in range(1, len(ORDER) + 1):
        allowed = set(ORDER[:k])
        print("prefix", k, file=sys.stderr, flush=True)

        def cut_call(name, *args, _allowed=allowed):
            if name in _allowed or name not in ORDER:
                return real_call(name, *args)

        ops._call = cut_call
        DenoiseGraph.STEPS_PER_GRAPH = 1
        dg = DenoiseGraph(model, n_poses, n_steps, src, q, False, dev)
        dg.run(T_seed.double().to(dev), src, q, rows, rows_all, None, 0)          # captures (+ runs n_steps replays)
        out["edges"] = int(dg.capacity)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dg.counter.zero_()
        e0.record()
        for i in range(replays):
            dg.graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / replays
        out[ORDER[k - 1]] = {"prefix_us": round(us, 2), "marginal_us": round(us - prev, 2)}
        prev = us
        ops._call = real_call
out["step_us"] = round(prev, 2)
# the full step unrolled k times per graph launch: what a launch costs on top of its kernels
with torch.no_grad():
    for spg in (1, 4, 16):
        DenoiseGraph.STEPS_PER_GRAPH = spg
        dg = DenoiseGraph(model, n_poses, n_steps, src, q, False, dev)
        dg.run(T_seed.double().to(dev), src, q, rows, rows_all, None, 0)
        torch.cuda.synchronize()
        g = dg.graph_multi if spg > 1 else dg.graph
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(max(1, replays // spg)):
            dg.counter.zero_() if False else None
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        out[f"step_us_{spg}_per_graph"] = round(1e3 * e0.elapsed_time(e1) / (max(1, replays // spg) * spg), 2)
print(json.dumps(out))
