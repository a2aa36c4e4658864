#!/usr/bin/env python
"""Benchmark of the patch-by-patch Generator inference path (BASELINE.json: output megapixels/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg1|cfg3|cfg4|cfg5band] [--precision fp16|bf16|fp32]
    python bench.py --impl reference ...        # the reference algorithm's CPU path (oracle port) on the host cores

One "step" = one Generator pass over one synthetic texture of the workload (random-init weights of the named
architecture, host-supplied z / noise grids).  `value` is timed with inputs resident in HBM (CUDA events around
each step, L2 flushed between steps); `e2e` goes through the public API (utils.sample_from_gen_PatchByPatch_test)
with pinned HOST noise in and a pinned HOST image out.  N > 1: the texture grows to N row bands of whole patch rows,
one band per rank, halo rows exchanged per layer over NCCL (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (ctor kwargs, patch rows, patch cols, description)      -- BASELINE.json `configs`
    "cfg1": (dict(z_dim=128, G_ch=52, n_layers_G=6, attention=True, leak=0.02, type_norm="BN", outer_padding="replicate"),
             9, 9, "241 Generator (n_layers_G=6, BN, attention, replicate), 1024x1024 -> 9x9 patch grid"),
    "cfg2": (dict(z_dim=128, G_ch=52, n_layers_G=6, attention=True, leak=0.02, type_norm="BN", outer_padding="replicate"),
             7, 21, "241 Generator (n_layers_G=6, BN, attention, replicate), 7x21 patch grid (896x2688)"),
    "cfg3": (dict(z_dim=128, G_ch=52, n_layers_G=5, attention=True, leak=0.02, type_norm="SSM", outer_padding="replicate"),
             61, 61, "34 Generator (n_layers_G=5, SSM, attention), 61x61 patch grid (3904x3904)"),
    "cfg4": (dict(z_dim=128, G_ch=52, n_layers_G=4, attention=True, leak=0.02, type_norm="BN", outer_padding="replicate"),
             129, 129, "417 Generator (n_layers_G=4, BN, attention), one 4096x4096 texture (129x129 patch grid)"),
    "cfg5band": (dict(z_dim=128, G_ch=52, n_layers_G=6, attention=True, leak=0.02, type_norm="BN", outer_padding="replicate"),
                 16, 129, "241 Generator, 16x129-patch row band (2048x16512) of a large seamless texture"),
}
L2_FLUSH_BYTES = 256 << 20


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def make_weights(kw, seed=1234):
    """Random-init weights of the named architecture with the reference's own init scheme (utils.py:745-762),
    realised by the drop-in Generator's constructor (orthogonal convs, zero biases, BN weight ~ N(1, .02))."""
    import infinite_texture_gans_b200 as itg
    torch.manual_seed(seed)
    net = itg.ResidualPatchGenerator(**kw)
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


def make_noise(kw, th, tw, seed=4321):
    import infinite_texture_gans_b200 as itg
    torch.manual_seed(seed)
    return itg.utils.draw_noise(1, kw["z_dim"], 4, kw["n_layers_G"], 1, kw["type_norm"], th, tw)


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference algorithm on the host cores (oracle port, Oracle B schedule)
# --------------------------------------------------------------------------------------------------
def cpu_reference_rate(kw, th, tw, sd, budget_s=20.0, max_steps=None):
    """Time the CPU restatement of utils.sample_from_gen_PatchByPatch_test (oracle/itg_oracle.py, the shipped
    3x3 sub-image schedule) on a bounded number of sub-image steps; returns (MP/s, threads, description)."""
    from oracle import itg_oracle as O
    ocfg = O.GenCfg(**kw)
    P = ocfg.patch_px
    H, W = th * P, tw * P
    z, maps = make_noise(kw, th, tw)
    geo = O.geometry(H, W, ocfg)
    total_steps = geo["steps_h"] * geo["steps_w"]
    threads = torch.get_num_threads()
    with torch.no_grad():
        O.sample_patch_by_patch(sd, ocfg, H, W, z, maps, max_steps=min(2, total_steps))     # warm-up (oneDNN primitives)
        t0 = time.perf_counter()
        O.sample_patch_by_patch(sd, ocfg, H, W, z, maps, max_steps=1)
        t1 = time.perf_counter() - t0
        n = total_steps if max_steps is None else min(max_steps, total_steps)
        n = max(1, min(n, int(budget_s / max(t1, 1e-3))))
        n = min(n, geo["steps_w"]) if n < total_steps else n      # partial runs stay inside the first row of sub-images
        t0 = time.perf_counter()
        O.sample_patch_by_patch(sd, ocfg, H, W, z, maps, max_steps=None if n == total_steps else n)
        dt = time.perf_counter() - t0
    full_time = dt * total_steps / n
    mp = H * W / 1e6
    sample = (f"{n} of {total_steps} sub-image steps (3x3 patches each) of the {th}x{tw} grid timed in {dt:.2f} s"
              + ("" if n == total_steps else ", extrapolated linearly in steps"))
    return mp / full_time, threads, sample, full_time


def run_reference(args, kw, th, tw, desc):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    sd = make_weights(kw)
    rates, sample, threads = [], "", 1
    for i in range(args.warmup + args.steps):
        r, threads, sample, _ = cpu_reference_rate(kw, th, tw, sd, budget_s=12.0)
        if i >= args.warmup:
            rates.append(r)
    v = sum(rates) / len(rates)
    P = 4 * 2 ** (kw["n_layers_G"] - 1)
    line = {"impl": "reference", "metric": "output megapixels/sec (Generator, local padding)", "value": v, "unit": "MP/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": th * tw * P * P / 1e6 / v * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "schedule": "sequential 3x3 sub-images (as shipped)"},
            "cpu_baseline": {"value": v, "unit": "MP/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# per-launch roofline of the conv kernel
# --------------------------------------------------------------------------------------------------
def launch_profile(plan, reps=5):
    """CUDA-event time of every launch of the plan (eager, same stream), median of `reps`."""
    n = plan.n_launches
    times = [[] for _ in range(n)]
    for _ in range(reps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        evs[0].record()
        for i, fn in enumerate(plan.fns):
            fn()
            evs[i + 1].record()
        torch.cuda.synchronize()
        for i in range(n):
            times[i].append(evs[i].elapsed_time(evs[i + 1]))
    return [sorted(t)[len(t) // 2] for t in times]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"], help="N > 1: halo rows over peer-mapped memory (NVLink P2P) or NCCL send/recv")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default="")
    args = ap.parse_args()
    kw, th, tw, desc = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, kw, th, tw, desc)
    args.warmup = max(args.warmup, 3)

    import infinite_texture_gans_b200 as itg
    from infinite_texture_gans_b200 import _lib as L
    from infinite_texture_gans_b200.config import GenConfig, flops_per_patch
    from infinite_texture_gans_b200.halo import BandHalo, P2PBandHalo

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # rank 0 prints exactly one JSON line on stdout: NCCL's version banner (written to fd 1 at communicator creation) goes to stderr
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
        # pinned host buffers should live on the GPU's own NUMA node (8 ranks copying 29 MB images per step share the host's memory system)
        try:
            import pynvml
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(local)
            try:
                h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
            except Exception:                                        # noqa: BLE001
                h = pynvml.nvmlDeviceGetHandleByIndex(local)
            pynvml.nvmlDeviceSetCpuAffinity(h)
        except Exception as e:                                       # noqa: BLE001  (affinity is an optimisation only)
            print(f"[bench] rank {rank}: could not set CPU affinity ({type(e).__name__}: {e})", file=sys.stderr)

    cfg = GenConfig(**kw)
    P = cfg.patch_px
    sd = make_weights(kw)
    net = itg.ResidualPatchGenerator(**kw, precision=args.precision)
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    eng = net.engine()

    # the texture: `world` row bands of th patch rows each; this rank owns band `rank`
    TH = th * world
    z_full, maps_full = make_noise(kw, TH, tw)
    b = cfg.base_res
    r0 = rank * th
    z_band = z_full[:, :, r0 * b:(r0 + th) * b + 2].contiguous()
    maps_band = None
    if maps_full is not None:
        maps_band = [m[:, :, r0 * b * 2 ** i:(r0 + th) * b * 2 ** i + 4].contiguous() for i, m in enumerate(maps_full)]
    plan = eng.plan(th, tw, L.IMG_MERGED)
    band, p2p = None, False
    if world > 1:
        if args.halo == "p2p":
            try:
                band, p2p = P2PBandHalo(plan), True
            except Exception as e:                                   # noqa: BLE001  (IPC unavailable: fall back to NCCL send/recv)
                print(f"[bench] rank {rank}: P2P halo setup failed ({type(e).__name__}: {e}); using NCCL send/recv", file=sys.stderr)
            ok = torch.tensor([1 if p2p else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                if p2p:
                    band.close()
                band, p2p = None, False
        if band is None:
            band = BandHalo()
    hooks = band.hooks(plan) if band is not None else None
    use_graph = (not args.no_graph) and world == 1

    z_pin = z_band.pin_memory()
    maps_pin = None if maps_band is None else [m.pin_memory() for m in maps_band]
    out_pin = torch.empty((1, cfg.img_ch, th * P, tw * P), dtype=torch.float32).pin_memory()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    band_graph = None

    def step_eager():
        if p2p:
            band.begin_step()
        plan.run(hooks)

    def step_device():
        if use_graph:
            eng.replay(th, tw, L.IMG_MERGED)
        elif band_graph is not None:
            band_graph.replay()
        else:
            step_eager()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (`value`) ----------------
    plan.set_inputs(z_pin[0], None if maps_pin is None else [m[0, 0] for m in maps_pin])
    for _ in range(args.warmup):
        step_device()
    barrier()
    if p2p and not args.no_graph:
        # launches + device-side halo exchanges of one step in ONE CUDA graph (no NCCL inside, nothing to deadlock on)
        band_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(band_graph):
            step_eager()
        for _ in range(3):
            step_device()
        barrier()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = eng.backend.launches
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()                                  # evict the previous step's tensors from the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_device()
        e1.record()
        evs.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b_) for a, b_ in evs)
    per_step = plan.n_launches + ((len(plan.halo_points) + 1) if p2p else 0)      # + device-side halo exchanges + step counter
    launches = (per_step * args.steps) if (use_graph or band_graph is not None) else (eng.backend.launches - launches0)
    clk = clocks.stop()
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    mp_step = world * th * tw * P * P / 1e6
    ms_per_step = dev_ms / args.steps
    value = mp_step / (ms_per_step / 1e3)

    # ---------------- end to end through the public API (`e2e`) ----------------
    # (1) synchronous: one sampler call per step, image copied to pinned host memory, host waits for it
    def step_e2e():
        if world == 1:
            img = itg.utils.sample_from_gen_PatchByPatch_test(
                net, z_dim=cfg.z_dim, output_resolution_height=th * P, output_resolution_width=tw * P,
                noise=(z_pin, maps_pin), return_on_device=True, graph=use_graph)
        else:
            plan.set_inputs(z_pin[0], None if maps_pin is None else [m[0, 0] for m in maps_pin])
            step_device()
            img = plan.out
        out_pin.copy_(img, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_sync_value = mp_step / (float(t.item()) / args.steps)

    # (2) streaming: the same per-step copies (z and maps host -> device, fp32 image device -> pinned host), with step k's
    # image crossing PCIe while step k+1 computes (utils.generate_textures / utils.HostOutputPipe); every copy of the K
    # steps, including the last image's, completes inside the timed region
    def run_stream(n):
        if world == 1:
            got = 0
            for _img in itg.utils.generate_textures(net, ((z_pin, maps_pin) for _ in range(n)), th * P, tw * P, graph=use_graph):
                got += 1
            assert got == n
        else:
            in_flight = []
            for _ in range(n):
                plan.set_inputs(z_pin[0], None if maps_pin is None else [m[0, 0] for m in maps_pin])
                step_device()
                in_flight.append(pipe.push(plan.out))
                if len(in_flight) == pipe.depth:
                    pipe.wait(in_flight.pop(0))
            for slot in in_flight:
                pipe.wait(slot)

    pipe = itg.utils.HostOutputPipe(tuple(plan.out.shape), dev) if world > 1 else None
    # (3) as (2) with the output stage of test_sample.py (img * 0.5 + 0.5 -> 8-bit, what save_image writes) done on the device:
    # reported as e2e.u8_value, N = 1 only; the headline e2e stays the fp32 image the reference's sampler returns
    e2e_u8 = None
    if world == 1:
        def run_stream_u8(n):
            for _img in itg.utils.generate_textures(net, ((z_pin, maps_pin) for _ in range(n)), th * P, tw * P, graph=use_graph, out_format="uint8"):
                pass
        run_stream_u8(3)
        barrier()
        t0 = time.perf_counter()
        run_stream_u8(args.steps)
        barrier()
        e2e_u8 = mp_step / ((time.perf_counter() - t0) / args.steps)
    run_stream(3)
    barrier()
    t0 = time.perf_counter()
    run_stream(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = mp_step / (float(t.item()) / args.steps)
    h2d = z_pin.numel() * 4 + (0 if maps_pin is None else sum(m.numel() * 4 for m in maps_pin))
    d2h = out_pin.numel() * 4

    # ---------------- roofline of the conv kernel (per-launch CUDA events, eager) ----------------
    bf16_peak, bf16_sus, hbm_peak, peak_kind = peaks()
    roof, cpu_base = None, None
    if rank == 0:
        times = launch_profile(plan)
        conv_ms = sum(tm for (kind, op), tm in zip(plan.ops, times) if kind in ("conv", "ssm"))
        all_ms = sum(times)
        total_flops = flops_per_patch(cfg) * th * tw
        att_flops = 0.0
        if cfg.attention:
            c, px = 2 * cfg.G_ch, cfg.level_res(3) ** 2
            att_flops = 2.0 * th * tw * (px * c * (c // 8) * 2 + px * c * (c // 2) + px * (c // 2) * c + px * (px // 4) * (c // 8)
                                         + px * (px // 4) * (c // 2))
        conv_alg_flops = total_flops - att_flops
        achieved = conv_alg_flops / (conv_ms / 1e3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(f"{args.workload}:{args.precision}")
        roof = {"bound": "tensor", "kernel": "conv_umma_kernel + conv_tile_kernel (all conv launches of one step)" if args.precision != "fp32"
                else "conv_direct_kernel", "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s",
                "frac": achieved / bf16_peak, "peak_kind": f"{peak_kind} bf16 burst", "traffic": traffic,
                "traffic_note": "DRAM bytes of all conv launches of one step (ncu, profiles/r01_summary.md); compulsory z in + fp32 image out = "
                                f"{(cfg.z_dim * (th * 4 + 2) * (tw * 4 + 2) + cfg.img_ch * th * tw * P * P) * 4} B",
                "conv_ms_per_step": conv_ms, "all_launches_ms_per_step": all_ms,
                "step_tflops": total_flops * world / (ms_per_step / 1e3) / 1e12,
                "step_frac_of_peak": total_flops * world / (ms_per_step / 1e3) / 1e12 / (bf16_peak * world)}
        if args.profile_out:
            rows = []
            for (kind, op), tm in zip(plan.ops, times):
                name = getattr(op, "name", kind) if kind in ("conv", "att", "ssm") else kind
                rows.append({"launch": name, "kind": kind, "ms": tm})
            json.dump({"workload": desc, "precision": args.precision, "launches": rows}, open(args.profile_out, "w"), indent=1)
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            v, threads, sample, _ = cpu_reference_rate(kw, th, tw, sd, budget_s=15.0)
            cpu_base = {"value": v, "unit": "MP/s", "cores": threads, "kind": "port", "sample": sample}

    if rank == 0:
        line = {"metric": "output megapixels/sec (Generator, local padding)", "value": value, "unit": "MP/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.precision],
                "data": "synthetic",
                "config": {"workload": desc + (f"; x{world} row bands, one per GPU, per-layer halo rows over " + ("NVLink P2P (itg_halo_exchange)" if p2p else "NCCL send/recv") if world > 1 else ""),
                           "weights": "random init (reference init scheme)", "l2": f"flushed between steps ({L2_FLUSH_BYTES >> 20} MiB write)",
                           "launch": "CUDA graph replay" if (use_graph or band_graph is not None) else "eager launches", "wall_ms_per_step_incl_flush": t_wall / args.steps * 1e3},
                "clocks": clk, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "mode": "streaming public API (utils.generate_textures): per step H2D of the noise from pinned memory + D2H of the fp32 image into "
                                "pinned memory, image k's D2H overlapped with pass k+1; wall clock over K steps incl. the last copy",
                        "sync_value": e2e_sync_value, "sync_mode": "one blocking sample_from_gen_PatchByPatch_test call + D2H per step",
                        "u8_value": e2e_u8, "u8_mode": "as value, with test_sample.py's 8-bit output stage on the device (d2h = 1/4 of d2h_bytes_per_step)"},
                "roofline": roof}
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line), flush=True)
    if p2p:
        barrier()
        band.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
