#!/usr/bin/env python
"""Where does the end-to-end step go?  Times, with CUDA events, the device->pinned-host copy of one cfg2 image alone and while
Generator passes run on another stream, and the Generator pass alone / while copies run."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import infinite_texture_gans_b200 as itg
from infinite_texture_gans_b200 import _lib as L

kw, th, tw, _ = bench.WORKLOADS["cfg2"]
dev = torch.device("cuda", 0)
net = itg.ResidualPatchGenerator(**kw, precision="fp16"); net.load_state_dict(bench.make_weights(kw)); net = net.to(dev).eval()
eng = net.engine(); plan = eng.plan(th, tw, L.IMG_MERGED)
z, _ = bench.make_noise(kw, th, tw); plan.set_inputs(z[0])
for _ in range(3): eng.replay(th, tw, L.IMG_MERGED)
torch.cuda.synchronize()
stage = torch.empty_like(plan.out); host = torch.empty(plan.out.shape).pin_memory()
cs = torch.cuda.Stream(dev)

def ev(): return torch.cuda.Event(enable_timing=True)

def t_copy(n=10):
    a, b = ev(), ev()
    with torch.cuda.stream(cs):
        a.record(cs)
        for _ in range(n): host.copy_(stage, non_blocking=True)
        b.record(cs)
    return a, b, n

def t_comp(n=10):
    a, b = ev(), ev(); cur = torch.cuda.current_stream()
    a.record(cur)
    for _ in range(n): eng.replay(th, tw, L.IMG_MERGED)
    b.record(cur)
    return a, b, n

a, b, n = t_copy(); torch.cuda.synchronize(); print(f"D2H alone      : {a.elapsed_time(b)/n:.3f} ms  ({stage.numel()*4/1e6:.1f} MB)")
a, b, n = t_comp(); torch.cuda.synchronize(); print(f"pass alone     : {a.elapsed_time(b)/n:.3f} ms")
c1 = t_comp(20); c2 = t_copy(10); torch.cuda.synchronize()
print(f"pass | D2H     : {c1[0].elapsed_time(c1[1])/c1[2]:.3f} ms   D2H | pass: {c2[0].elapsed_time(c2[1])/c2[2]:.3f} ms")
# host-side cost of the pieces of one streaming iteration
t0 = time.perf_counter()
for _ in range(20): plan.set_inputs(z[0].pin_memory() if False else z[0])
torch.cuda.synchronize(); print(f"set_inputs (pageable z) host+dev: {(time.perf_counter()-t0)/20*1e3:.3f} ms")
zp = z.pin_memory()
t0 = time.perf_counter()
for _ in range(20): plan.set_inputs(zp[0])
torch.cuda.synchronize(); print(f"set_inputs (pinned z)   host+dev: {(time.perf_counter()-t0)/20*1e3:.3f} ms")
# the streaming loop itself, with per-iteration wall times
pipe = itg.utils.HostOutputPipe(tuple(plan.out.shape), dev)
prev = None; ts = []
for k in range(12):
    t0 = time.perf_counter()
    plan.set_inputs(zp[0]); eng.replay(th, tw, L.IMG_MERGED); slot = pipe.push(plan.out)
    t1 = time.perf_counter()
    if prev is not None: pipe.wait(prev)
    prev = slot; ts.append((t1 - t0, time.perf_counter() - t1))
pipe.wait(prev)
print("stream loop: enqueue ms / wait ms per iteration:", " ".join(f"{a*1e3:.2f}/{b*1e3:.2f}" for a, b in ts))
