"""In-graph timeline of the C2 forward (10k-pt scene, 128 poses): %globaltimer stamps (dedf_stamp) captured INTO the CUDA graph between the
kernels of both streams, read after a replay.   python profiles/run_timeline.py [n_poses=128]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusion_edf_b200 import FeaturedPoints, MultiscaleScoreModel, ops
from diffusion_edf_b200.synthetic import make_poses, make_scene, model_kwargs

n_poses = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MultiscaleScoreModel(**model_kwargs(), deterministic=True).eval().to(dev)
x, rgb = make_scene(10_000, seed=0)
Ts, t = make_poses(n_poses, x, seed=0)
key = FeaturedPoints(x.to(dev), rgb.to(dev), torch.zeros(len(x), dtype=torch.long, device=dev))
grasp = FeaturedPoints(torch.zeros(8, 3, device=dev), torch.zeros(8, 3, device=dev), torch.zeros(8, dtype=torch.long, device=dev))
Ts, t = Ts.to(dev), t.to(dev)
ops.TIMELINE = {"buf": torch.zeros(512, dtype=torch.int64, device=dev), "names": []}
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
with torch.no_grad():
    model(Ts, t, key, grasp)                       # plan + capture (the stamps are captured with the kernels)
    names = list(ops.TIMELINE["names"])
    n_per = None
    for _ in range(3):
        flush.fill_(1)
        model(Ts, t, key, grasp)
    torch.cuda.synchronize()
buf = ops.TIMELINE["buf"].cpu().tolist()
# the eager planning pass and the capture pass both appended names: the replay rewrites the slots of the LAST pass
starts = [i for i, (n, _) in enumerate(names) if n == "forward start"]
lo = starts[-1]
rows = [(buf[i] - buf[lo], names[i][0], names[i][1]) for i in range(lo, len(names))]
main = names[lo][1]
out = []
for ns, nm, st in sorted(rows):
    out.append({"us": round(ns / 1e3, 1), "stream": "main" if st == main else "geometry", "what": nm})
    print(f"{ns / 1e3:9.1f} us  {'main' if st == main else 'geo ':4s}  {nm}")
print(json.dumps({"n_poses": n_poses, "timeline": out}))
