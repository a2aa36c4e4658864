#!/usr/bin/env python
"""Multi-GPU parity check of the row-band split through the package's public sampler (run under torchrun, one rank per GPU):
every rank computes its band with bands.RowBandSampler (P2P halo exchange: eagerly, then from a CUDA graph; or torch.distributed
send/recv); rank 0 gathers the bands, computes the whole texture on one GPU and compares.  Also checks bands.sample_row_bands with the
device-side noise generator.  Exit code 0 = the bands match the single-GPU result bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/band_check.py [p2p|dist] [cfg2|cfg3|...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
import infinite_texture_gans_b200 as itg
from infinite_texture_gans_b200 import bands

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "p2p"
wl = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
kw, th, tw, desc = bench.WORKLOADS[wl]
th, tw = min(th, 7), min(tw, 21)
net = itg.ResidualPatchGenerator(**kw, precision="fp16")
net.load_state_dict(bench.make_weights(kw))
net = net.to(dev).eval()
TH = th * world + (1 if world > 1 else 0)                      # unequal bands: the first rank gets one patch row more
z, maps = bench.make_noise(kw, TH, tw)
ok = True
outs = {}
for graph in ((False, True) if mode == "p2p" else (False,)):
    s = bands.RowBandSampler(net, TH, tw, halo="p2p" if mode == "p2p" else "dist", graph=graph)
    s.set_noise(z, maps)
    for _ in range(3):
        band = s.step()
    torch.cuda.synchronize()
    full = s.gather(band.clone())
    if rank == 0:
        outs[graph] = full
    s.close()
if rank == 0:
    ref = itg.utils.generate_full_grid(net, z, maps)
    err = max((o - ref).abs().max().item() for o in outs.values())
    graph_ok = len(outs) == 1 or torch.equal(outs[False], outs[True])
    print(f"[band_check] mode={mode} world={world} {wl} ({TH}x{tw} patches): max|bands - single GPU| = {err:.3e}; graph replay identical: {graph_ok}", flush=True)
    ok = err <= 1e-6 and graph_ok
# device-side noise: every rank generates only its own band of the counter-based field; rank 0 compares with the one-GPU image of the whole field
H, W = TH * net.cfg.patch_px - 5, tw * net.cfg.patch_px - 9
img = bands.sample_row_bands(net, H, W, seed=2024, halo="p2p" if mode == "p2p" else "dist")
if rank == 0:
    geo = itg.utils.patch_grid_geometry(H, W, net.n_layers_G, net.cfg.base_res)
    zf, mf = itg.utils.draw_noise_device(net.cfg, geo["total_h"], geo["total_w"], 2024, device=dev)
    ref = itg.utils.generate_full_grid(net, zf.unsqueeze(0), None if mf is None else [m[None, None] for m in mf])[:, :, :H, :W]
    e2 = (img - ref).abs().max().item()
    print(f"[band_check] sample_row_bands with device noise ({geo['total_h']}x{geo['total_w']} patches, cropped to {H}x{W}): max|.| = {e2:.3e}", flush=True)
    ok = ok and e2 <= 1e-6
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
