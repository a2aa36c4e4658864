#!/usr/bin/env python
"""BASELINE.json config 4: a batch of independent 4096 x 4096 textures (417 Generator, n_layers_G=4; 129 x 129 patches each), data-parallel
over the GPUs of one box: replicas only, no collective on the data path (bands.generate_textures_replicas).  Run under torchrun, one rank
per GPU (or plain `python` for one GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/run_cfg4.py [--textures 256]

Every texture's noise is uploaded from pinned host memory and its 8-bit image (test_sample.py's output stage) lands in pinned host memory, copies
overlapped with compute; the time is wall clock over the whole batch, max over ranks.  Prints one JSON line on rank 0."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import infinite_texture_gans_b200 as itg
from infinite_texture_gans_b200 import bands
from infinite_texture_gans_b200.config import flops_per_patch

ap = argparse.ArgumentParser()
ap.add_argument("--textures", type=int, default=256)
ap.add_argument("--size", type=int, default=4096)
ap.add_argument("--out-format", default="uint8", choices=["uint8", "float32"])
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
kw, _, _, _ = bench.WORKLOADS["cfg4"]
net = itg.ResidualPatchGenerator(**kw, precision="fp16")
net.load_state_dict(bench.make_weights(kw))
net = net.to(dev).eval()
geo = itg.utils.patch_grid_geometry(a.size, a.size, net.n_layers_G, net.cfg.base_res)
th, tw = geo["total_h"], geo["total_w"]
# two pinned noise grids per rank, reused round-robin (256 distinct 137 MB grids would only measure the host's RNG)
g = torch.Generator().manual_seed(1000 + rank)
pool = [(torch.randn(1, net.cfg.z_dim, th * 4 + 2, tw * 4 + 2, generator=g).pin_memory(), None) for _ in range(2)]
noises = (pool[i % 2] for i in range(a.textures))
mine = len(range(rank, a.textures, world))
for _ in bands.generate_textures_replicas(net, (pool[i % 2] for i in range(2 * world)), a.size, a.size, rank=rank, world=world, out_format=a.out_format):
    pass                                                                                   # warm-up: plan, graph capture, staging buffers
torch.cuda.synchronize()
if dist is not None:
    dist.barrier()
t0 = time.perf_counter()
n, checksum = 0, 0
for idx, img in bands.generate_textures_replicas(net, noises, a.size, a.size, rank=rank, world=world, out_format=a.out_format):
    n += 1
    checksum += int(img[0, 0, 0]) if a.out_format == "uint8" else 0
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
cnt = torch.tensor([n], dtype=torch.int64, device=dev)
if dist is not None:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dist.all_reduce(cnt)
assert n == mine
if rank == 0:
    total = int(cnt.item())
    mp = total * a.size * a.size / 1e6
    fl = flops_per_patch(net.cfg) * th * tw * total
    s = float(dt.item())
    print(json.dumps({"workload": f"417 Generator, {total} independent {a.size}x{a.size} textures ({th}x{tw} patches each), {world} replica(s), "
                                  f"{mine} textures on rank 0; noise H2D + {a.out_format} image D2H per texture, overlapped",
                      "n_gpus": world, "seconds": s, "value": mp / s, "unit": "MP/s", "ms_per_texture_per_gpu": s / mine * 1e3,
                      "algorithmic_tflops": fl / s / 1e12, "frac_of_measured_bf16_peak": fl / s / 1e12 / (bench.peaks()[0] * world), "dtype": "f16"}), flush=True)
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
