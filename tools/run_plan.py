#!/usr/bin/env python
"""Run the launch plan of a bench workload a few times, eagerly (the command ncu captures)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--precision", default="fp16")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
import infinite_texture_gans_b200 as itg
from infinite_texture_gans_b200 import _lib as L
kw, th, tw, desc = bench.WORKLOADS[a.workload]
net = itg.ResidualPatchGenerator(**kw, precision=a.precision)
net.load_state_dict(bench.make_weights(kw))
net = net.cuda().eval()
z, maps = bench.make_noise(kw, th, tw)
plan = net.engine().plan(th, tw, L.IMG_MERGED)
plan.set_inputs(z[0], None if maps is None else [m[0, 0] for m in maps])
for _ in range(a.reps):
    plan.run()
torch.cuda.synchronize()
print("ok", plan.n_launches, "launches per pass;", [getattr(op, "name", k) for k, op in plan.ops])
