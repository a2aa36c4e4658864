#!/usr/bin/env python
"""Multi-GPU parity check of the row-band split (run under torchrun, one rank per GPU):
every rank computes its band with the P2P halo exchange (eager, then from a CUDA graph); rank 0 also computes the whole
texture on one GPU and compares.  Exit code 0 = bands match the single-GPU result."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
import infinite_texture_gans_b200 as itg
from infinite_texture_gans_b200 import _lib as L
from infinite_texture_gans_b200.halo import P2PBandHalo, BandHalo

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "p2p"
wl = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
kw, th, tw, desc = bench.WORKLOADS[wl]
th = min(th, 7)
net = itg.ResidualPatchGenerator(**kw, precision="fp16")
net.load_state_dict(bench.make_weights(kw))
net = net.to(dev).eval()
eng = net.engine()
cfg, b = net.cfg, net.cfg.base_res
z, maps = bench.make_noise(kw, th * world, tw)
r0 = rank * th
zb = z[0, :, r0 * b:(r0 + th) * b + 2].contiguous()
mb = None if maps is None else [m[0, 0, r0 * b * 2 ** i:(r0 + th) * b * 2 ** i + 4].contiguous() for i, m in enumerate(maps)]
plan = eng.plan(th, tw, L.IMG_MERGED)
plan.set_inputs(zb, mb)
band = P2PBandHalo(plan) if mode == "p2p" else BandHalo()
hooks = band.hooks(plan)

def step():
    if mode == "p2p":
        band.begin_step()
    plan.run(hooks)

for _ in range(3):
    step()
torch.cuda.synchronize()
eager = plan.out.clone()
graph_ok = True
if mode == "p2p":
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    graph_ok = torch.equal(plan.out, eager)
outs = [torch.empty_like(eager) for _ in range(world)]
dist.all_gather(outs, eager)
ok = True
if rank == 0:
    full = itg.utils.generate_full_grid(net, z, maps)
    got = torch.cat(outs, dim=2)
    err = (got - full).abs().max().item()
    print(f"[band_check] mode={mode} world={world} {wl}: max|bands - single GPU| = {err:.3e}; graph replay identical: {graph_ok}", flush=True)
    ok = err <= 1e-6 and graph_ok
flag = torch.tensor([1 if (ok and graph_ok) else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if mode == "p2p":
    band.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
