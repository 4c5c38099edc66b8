"""BASELINE config C5: trainer step (forward + backward + Adam) of the panda_bottle/pick_lowres score model (identical kwargs
to panda_mug) on a synthetic demo batch: 6000-point scene, nT = 20 poses (2 time schedules x 10 reference
points), Adam(lr 3e-4, betas (0.9, 0.98), eps 1e-9, wd 1e-4, amsgrad) (train_configs.yaml:70-75).  Under torchrun every
rank trains on its own synthetic demo (seed = rank) and gradients are averaged with ONE bucketed all-reduce per step.

    python profiles/run_c5.py [steps=10]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 profiles/run_c5.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from diffusion_edf_b200 import FeaturedPoints, MultiscaleScoreModel, ops, parallel
from diffusion_edf_b200.synthetic import make_poses, make_scene, model_kwargs

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = MultiscaleScoreModel(**model_kwargs(), deterministic=True).train().to(dev)
opt = torch.optim.Adam(model.parameters(), lr=3e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-4, amsgrad=True)
x, rgb = make_scene(6000, seed=rank)
Ts, t = make_poses(20, x, seed=rank, spread=4.0)
g = torch.Generator().manual_seed(rank)
ta, tl = torch.randn(20, 3, generator=g), torch.randn(20, 3, generator=g)
d = lambda v: v.to(dev)
args = (d(Ts), d(t), FeaturedPoints(d(x), d(rgb), torch.zeros(len(x), dtype=torch.long, device=dev)),
        FeaturedPoints(torch.zeros(800, 3, device=dev), torch.zeros(800, 3, device=dev), torch.zeros(800, dtype=torch.long, device=dev)), d(ta), d(tl))


def step():
    opt.zero_grad(set_to_none=True)
    loss, *_ = model.get_train_loss(*args)
    loss.backward()
    n = parallel.allreduce_gradients(model.parameters())
    opt.step()
    return loss, n


losses = []
for _ in range(3):
    losses.append(float(step()[0].detach()))
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
k0 = ops.LAUNCHES
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(steps):
    loss, n_coll = step()
    losses.append(float(loss.detach()))
e1.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
tt = torch.tensor([e0.elapsed_time(e1) / steps, 1e3 * wall / steps], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": "C5: trainer step fwd+bwd+Adam, 6000-pt synthetic scene, 20 poses per rank", "n_gpus": world,
                      "ms_per_step_device": float(tt[0]), "ms_per_step_wall": float(tt[1]), "steps": steps,
                      "demos_per_s": world / (float(tt[1]) / 1e3), "dedf_kernel_launches_per_step": (ops.LAUNCHES - k0) / steps,
                      "grad_allreduces_per_step": n_coll, "loss_first": losses[0], "loss_last": losses[-1],
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))
if world > 1:
    dist.destroy_process_group()
