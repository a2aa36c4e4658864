#!/usr/bin/env python
"""BASELINE.json config 5: one seamless 65536 x 65536 texture (241 Generator, 513 x 513 patches of 128 px) split into row
bands over the GPUs of one box, halo rows over NVLink P2P.  Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 tools/run_cfg5.py [--rows 513 --cols 513]

Prints one JSON line (rank 0): MP/s, ms per pass (CUDA events, max over ranks), algorithmic TFLOP/s and fraction of the
measured bf16 peak, and a seam check: the patches on both sides of the first band boundary are compared with a
single-GPU one-shot computation of a small window around the seam (same z), which must agree to float rounding."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
import infinite_texture_gans_b200 as itg
from infinite_texture_gans_b200 import _lib as L
from infinite_texture_gans_b200.config import flops_per_patch
from infinite_texture_gans_b200.halo import P2PBandHalo

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=513)
ap.add_argument("--cols", type=int, default=513)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--device-noise", action="store_true", help="draw every band's latent grid on its GPU (counter-based Philox field, itg_noise_normal) instead of on the host RNG")
a = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
kw, _, _, _ = bench.WORKLOADS["cfg2"]
net = itg.ResidualPatchGenerator(**kw, precision="fp16")
net.load_state_dict(bench.make_weights(kw))
net = net.to(dev).eval()
eng, cfg = net.engine(), net.cfg
b, P = cfg.base_res, cfg.patch_px

# bands of whole patch rows, as equal as possible
base, extra = divmod(a.rows, world)
heights = [base + (1 if r < extra else 0) for r in range(world)]
starts = [sum(heights[:r]) for r in range(world)]
r0, th, tw = starts[rank], heights[rank], a.cols

# every rank draws the same full latent grid (same seed) and keeps its band (+ the 1-px ring rows it shares with its neighbours)
t0 = time.time()
g = torch.Generator().manual_seed(4321)
full_rows = a.rows * b + 2
chunk = 64                                              # rows of the full grid drawn per call: bounded host memory
if a.device_noise:
    zb = itg.utils.draw_noise_device(cfg, a.rows, a.cols, 4321, rows=(r0, r0 + th), device=dev)[0]
    torch.cuda.synchronize()
else:
    zb = torch.empty((cfg.z_dim, th * b + 2, tw * b + 2), dtype=torch.float32)
for y0 in range(0, full_rows if not a.device_noise else 0, chunk):
    n = min(chunk, full_rows - y0)
    blk = torch.randn(cfg.z_dim, n, tw * b + 2, generator=g)
    lo, hi = max(y0, r0 * b), min(y0 + n, (r0 + th) * b + 2)
    if lo < hi:
        zb[:, lo - r0 * b:hi - r0 * b] = blk[:, lo - y0:hi - y0]
t_noise = time.time() - t0

plan = eng.plan(th, tw, L.IMG_MERGED)
plan.set_inputs(zb)
band = P2PBandHalo(plan)
hooks = band.hooks(plan)


def step():
    band.begin_step()
    plan.run(hooks)


for _ in range(a.warmup):
    step()
torch.cuda.synchronize()
dist.barrier()
evs = []
for _ in range(a.steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    evs.append((e0, e1))
torch.cuda.synchronize()
dist.barrier()
ms = torch.tensor([sum(x.elapsed_time(y) for x, y in evs) / a.steps], dtype=torch.float64, device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms = float(ms.item())

# ---- seam check at the boundary between band 0 and band 1 ----
ncol = min(12, tw)
seam = None
if world > 1:
    mine = None
    if rank == 0:
        mine = plan.out[:, :, (th - 1) * P:th * P, :ncol * P].contiguous()       # last patch row of band 0
    if rank == 1:
        mine = plan.out[:, :, :P, :ncol * P].contiguous()                         # first patch row of band 1
    bufs = [torch.empty((1, cfg.img_ch, P, ncol * P), dtype=torch.float32, device=dev) for _ in range(2)]
    if rank == 1:
        dist.send(mine, dst=0)
    if rank == 0:
        bufs[0] = mine
        dist.recv(bufs[1], src=1)
        # window: 3 patch rows above and below the seam, ncol + 2 columns, computed one-shot on this GPU from the same z
        wr0 = heights[0] - 3
        gz = torch.Generator().manual_seed(4321)
        zw = torch.empty((1, cfg.z_dim, 6 * b + 2, (ncol + 2) * b + 2))
        if a.device_noise:
            zw = itg.utils.draw_noise_device(cfg, a.rows, a.cols, 4321, rows=(wr0, wr0 + 6), device=dev)[0][:, :, :(ncol + 2) * b + 2].unsqueeze(0).contiguous()
        for y0 in range(0, full_rows if not a.device_noise else 0, chunk):
            n = min(chunk, full_rows - y0)
            blk = torch.randn(cfg.z_dim, n, tw * b + 2, generator=gz)
            lo, hi = max(y0, wr0 * b), min(y0 + n, (wr0 + 6) * b + 2)
            if lo < hi:
                zw[0, :, lo - wr0 * b:hi - wr0 * b] = blk[:, lo - y0:hi - y0, :(ncol + 2) * b + 2]
            if y0 + n >= (wr0 + 6) * b + 2:
                break
        win = itg.utils.generate_full_grid(net, zw).clone()                        # (1, 3, 6P, (ncol+2)P)
        # compare columns 2 .. ncol-1 (two patches away from the window's artificial left / right border)
        ref_top = win[:, :, 2 * P:3 * P, 2 * P:ncol * P]
        ref_bot = win[:, :, 3 * P:4 * P, 2 * P:ncol * P]
        seam = max((bufs[0][:, :, :, 2 * P:] - ref_top).abs().max().item(), (bufs[1][:, :, :, 2 * P:] - ref_bot).abs().max().item())

if rank == 0:
    peak = bench.peaks()[0]
    mp = a.rows * P * a.cols * P / 1e6
    fl = flops_per_patch(cfg) * a.rows * a.cols
    print(json.dumps({"workload": f"241 Generator, one seamless {a.rows * P} x {a.cols * P} texture ({a.rows} x {a.cols} patches), "
                                  f"{world} row bands, halo rows over NVLink P2P", "n_gpus": world, "ms_per_pass": ms, "value": mp / (ms / 1e3),
                      "unit": "MP/s", "algorithmic_tflops": fl / (ms / 1e3) / 1e12, "frac_of_measured_bf16_peak": fl / (ms / 1e3) / 1e12 / (peak * world),
                      "band_rows": heights, "arena_gb_per_gpu": plan.arena_bytes / 1e9, "halo_bytes_pushed_per_pass_rank0": band.bytes_per_step,
                      "noise_draw_s": t_noise, "seam_max_abs_vs_single_gpu_window": seam, "dtype": "f16"}), flush=True)
dist.barrier()
band.close()
dist.destroy_process_group()
