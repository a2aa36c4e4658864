#!/usr/bin/env python
"""Benchmark of the patch-by-patch Generator inference path (BASELINE.json: output megapixels/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg1|cfg2|cfg3|cfg4|cfg5band] [--precision fp16|fp32]
    python bench.py --impl reference ...        # the reference algorithm's CPU path (oracle port) on the host cores

One "step" = one Generator pass over one synthetic texture of the workload (random-init weights of the named
architecture, host-supplied z / noise grids).  Without --workload the headline is BASELINE.json's largest single-GPU
configuration (cfg3: the SSM Generator on a 61x61 patch grid) and the line carries the 7x21 grid (cfg2) and the
65536^2 texture's row band (cfg5band) in `extra`, each with its own roofline, e2e and parity.  `value` is timed with
inputs resident in HBM (CUDA events around each step, L2 flushed between steps); `e2e` goes through the public
streaming API with pinned HOST noise in and a pinned HOST image out; `parity` is the timed output's max-abs error
against the CPU oracle on a window.  N > 1: the texture grows to N row bands of that size, one band per rank
(bands.RowBandSampler), halo rows exchanged per layer over NVLink P2P (weak scaling).
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
HEADLINE = "cfg3"                  # the largest single-GPU configuration of BASELINE.json (20.4 algorithmic TFLOP per pass)
EXTRA = ("cfg2", "cfg5band")       # reported in the same JSON line (`extra`), each with its own roofline / e2e / parity


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
# per-launch timing and roofline
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


def attention_flops(cfg, th, tw):
    if not cfg.attention:
        return 0.0
    c, px = 2 * cfg.G_ch, cfg.level_res(3) ** 2
    return 2.0 * th * tw * (px * c * (c // 8) * 2 + px * c * (c // 2) + px * (c // 2) * c + px * (px // 4) * (c // 8) + px * (px // 4) * (c // 2))


def ssm_flops(cfg, th, tw):
    """Algorithmic FLOPs of the StochasticSpatialModulation convs (mlp_shared 1 -> 128 on the (r+2)^2 map, embed 128 -> 2C): layers.py:228-230."""
    if cfg.type_norm != "SSM":
        return 0.0
    fl = 0.0
    for k, (ci, co) in enumerate(cfg.block_channels(), start=1):
        r = cfg.level_res(k)
        px, px2 = th * r * tw * r, (th * r + 2) * (tw * r + 2)
        for c in ([ci, co] + ([ci] if ci != co else [])):
            fl += 2.0 * 9 * (px2 * 128 + px * 128 * 2 * c)
    return fl


def run_workload(name, args, ctx, steps, warmup, *, with_cpu, with_parity, profile_out=""):
    """Time one BASELINE.json workload on this rank's GPU (N > 1: one row band of that size per rank) and return its result entry."""
    import infinite_texture_gans_b200 as itg
    from infinite_texture_gans_b200 import _lib as L
    from infinite_texture_gans_b200 import bands
    from infinite_texture_gans_b200.config import GenConfig, flops_per_patch
    rank, world, dev, dist = ctx["rank"], ctx["world"], ctx["dev"], ctx["dist"]
    kw, th, tw, desc = WORKLOADS[name]
    cfg = GenConfig(**kw)
    P = cfg.patch_px
    sd = make_weights(kw)
    net = itg.ResidualPatchGenerator(**kw, precision=args.precision)
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    eng = net.engine()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # the texture: `world` row bands of th patch rows each; this rank owns band `rank` (bands.RowBandSampler, the package's multi-GPU sampler)
    TH = th * world
    z_full, maps_full = make_noise(kw, TH, tw)
    sampler = None
    if world > 1:
        sampler = bands.RowBandSampler(net, TH, tw, halo="auto" if args.halo == "p2p" else "dist", graph=not args.no_graph)
        plan, r0 = sampler.plan, sampler.rows[0]
        assert sampler.rows[1] - r0 == th
    else:
        plan, r0 = eng.plan(th, tw, L.IMG_MERGED), 0
    z_band, maps_band = bands.band_noise(cfg, z_full, maps_full, r0, r0 + th)
    z_pin = z_band.pin_memory()
    maps_pin = None if maps_band is None else [m.pin_memory() for m in maps_band]
    use_graph = (not args.no_graph) and world == 1
    flush = ctx["flush"]

    def step_device():
        if sampler is not None:
            sampler.step()
        elif use_graph:
            eng.replay(th, tw, L.IMG_MERGED)
        else:
            plan.run()

    # ---------------- device-resident timing (`value`) ----------------
    plan.set_inputs(z_pin, maps_pin)
    for _ in range(max(warmup, 3)):
        step_device()
    barrier()
    clocks = ClockSampler(ctx["local"])
    clocks.start()
    launches0 = eng.backend.launches
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(steps):
        flush.zero_()                                  # evict the previous step's tensors from the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_device()
        e1.record()
        evs.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b_) for a, b_ in evs)
    graphed = use_graph or (sampler is not None and sampler.use_graph)
    per_step = sampler.launches_per_step if sampler is not None else plan.n_launches
    launches = (per_step * steps) if graphed else (eng.backend.launches - launches0)
    clk = clocks.stop()
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    mp_step = world * th * tw * P * P / 1e6
    ms_per_step = dev_ms / steps
    value = mp_step / (ms_per_step / 1e3)

    # ---------------- end to end through the public API (`e2e`) ----------------
    def timed(fn, n):
        fn(3)
        barrier()
        t0 = time.perf_counter()
        fn(n)
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return mp_step / (float(tt.item()) / n)

    e2e_steps = max(3, min(steps, 20))

    def run_stream(fmt):
        def go(n):
            noises = ((z_pin, maps_pin) for _ in range(n))
            if sampler is not None:
                it = sampler.generate(noises, out_format=fmt)                  # bands.RowBandSampler.generate: this rank's band, streamed
            else:
                it = itg.utils.generate_textures(net, noises, th * P, tw * P, graph=use_graph, out_format=fmt)
            got = sum(1 for _ in it)
            assert got == n
        return go

    e2e_value = timed(run_stream("float32"), e2e_steps)
    e2e_u8 = timed(run_stream("uint8"), e2e_steps)
    e2e_sync = None
    if world == 1:
        out_pin = torch.empty((1, cfg.img_ch, th * P, tw * P), dtype=torch.float32).pin_memory()

        def run_sync(n):                              # one blocking call of the reference-signature sampler + D2H per step
            for _ in range(n):
                img = itg.utils.sample_from_gen_PatchByPatch_test(net, z_dim=cfg.z_dim, output_resolution_height=th * P, output_resolution_width=tw * P,
                                                                  noise=(z_pin.unsqueeze(0), None if maps_pin is None else [m[None, None] for m in maps_pin]),
                                                                  return_on_device=True, graph=use_graph)
                out_pin.copy_(img, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        e2e_sync = timed(run_sync, e2e_steps)
    h2d = z_pin.numel() * 4 + (0 if maps_pin is None else sum(m.numel() * 4 for m in maps_pin))
    d2h = cfg.img_ch * th * P * tw * P * 4

    # ---------------- roofline (per-launch CUDA events, eager) / parity / cpu baseline: rank 0 ----------------
    bf16_peak, bf16_sus, hbm_peak, peak_kind = peaks()
    roof, cpu_base, parity = None, None, None
    plan.set_inputs(z_pin, maps_pin)
    if sampler is None:
        plan.run()
    else:
        sampler.step()
    torch.cuda.synchronize()
    # N > 1: the first patch row BELOW the band-0 / band-1 seam (rank 1's output, its top halo rows came from rank 0 over NVLink) against the
    # CPU oracle on a noise window that straddles the seam (three patch rows of each band, two-patch margin) -- checker use of oracle/
    seam = None
    if world > 1 and with_parity:
        seam_t = torch.full((1,), -1.0, device=dev)
        if rank == 1:
            from oracle import itg_oracle as O
            from oracle import window as OW
            ocfg = O.GenCfg(**kw)
            cols = min(5, tw)
            zc, mc = OW.crop_noise(ocfg, z_full, maps_full, r0 - 3, r0 + 3, 0, cols)
            with torch.no_grad():
                ref = O.forward_merged(sd, ocfg, zc, mc)
            lc = cols if cols == tw else cols - 2
            got = plan.out[:, :, :P, :lc * P].detach().float().cpu()
            seam_t[0] = (got - ref[:, :, 3 * P:4 * P, :lc * P]).abs().max().item()
        dist.all_reduce(seam_t, op=dist.ReduceOp.MAX)
        tol = 2e-2 if args.precision != "fp32" else 1e-3
        seam = {"max_abs": seam_t.item(), "tolerance": tol, "ok": 0.0 <= seam_t.item() <= tol,
                "what": f"patch row {r0 if rank == 1 else th} (first row of band 1) x patch columns [0, {min(5, tw) if min(5, tw) == tw else min(5, tw) - 2}) against the oracle on the "
                        "noise window of patch rows [seam - 3, seam + 3)"}
    if rank == 0:
        total_flops = flops_per_patch(cfg) * th * tw
        if world == 1:
            times = launch_profile(plan)
            fams = {}
            for (kind, op), tm in zip(plan.ops, times):
                fams[kind] = fams.get(kind, 0.0) + tm
            conv_ms = fams.get("conv", 0.0) + fams.get("ssm", 0.0)
            conv_alg = total_flops - attention_flops(cfg, th, tw)
            achieved = conv_alg / (conv_ms / 1e3) / 1e12
            by = {}
            sf = ssm_flops(cfg, th, tw) if plan.fuse_ssm else 0.0
            if sf:
                by["ssm_fused2_kernel (StochasticSpatialModulation, cta_group::2)"] = {"ms": fams["ssm"], "tflops": sf / (fams["ssm"] / 1e3) / 1e12,
                                                                                      "frac": sf / (fams["ssm"] / 1e3) / 1e12 / bf16_peak}
            by["conv_pair_kernel + conv_tile_kernel + conv_umma_kernel (conv2d_lp / 1x1 launches)"] = {
                "ms": fams.get("conv", 0.0), "tflops": (conv_alg - sf) / (max(fams.get("conv", 0.0), 1e-9) / 1e3) / 1e12,
                "frac": (conv_alg - sf) / (max(fams.get("conv", 0.0), 1e-9) / 1e3) / 1e12 / bf16_peak}
            traffic = None
            tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
            if os.path.exists(tp):
                traffic = json.load(open(tp)).get(f"{name}:{args.precision}")
            roof = {"bound": "tensor", "kernel": "all conv-stack launches of one step (conv2d_lp, 1x1 shortcuts, SSM mlp_shared + embed); per kernel family in by_kernel",
                    "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s", "frac": achieved / bf16_peak, "peak_kind": f"{peak_kind} bf16 burst",
                    "frac_of_sustained": achieved / bf16_sus, "traffic": traffic,
                    "traffic_note": "DRAM bytes (read + write) of the conv-stack launches of one step from the committed ncu capture (profiles/r02_*), null if none was taken "
                                    f"for this workload; compulsory noise in + fp32 image out = {h2d + d2h} B",
                    "conv_ms_per_step": conv_ms, "all_launches_ms_per_step": sum(times), "by_kernel": by,
                    "algorithmic_flops_per_step": conv_alg}
            if profile_out:
                rows = []
                for (kind, op), tm in zip(plan.ops, times):
                    nm = getattr(op, "name", kind) if kind in ("conv", "att", "ssm") else kind
                    rows.append({"launch": nm, "kind": kind, "ms": tm})
                json.dump({"workload": desc, "precision": args.precision, "launches": rows}, open(profile_out, "w"), indent=1)
        else:
            roof = {"bound": "tensor", "kernel": "whole step (per-launch events are taken at N = 1)", "achieved": total_flops * world / (ms_per_step / 1e3) / 1e12 / world,
                    "peak": bf16_peak, "unit": "TFLOP/s per GPU", "frac": total_flops / (ms_per_step / 1e3) / 1e12 / bf16_peak, "peak_kind": f"{peak_kind} bf16 burst", "traffic": None}
        roof["step_tflops"] = total_flops * world / (ms_per_step / 1e3) / 1e12
        roof["step_frac_of_peak"] = roof["step_tflops"] / (bf16_peak * world)
        if with_parity:
            # the timed configuration's own output against the CPU oracle (checker use of oracle/): top-left 5 x 5 patch window of rank 0's band,
            # two-patch margin towards the interior (oracle/window.py)
            from oracle import itg_oracle as O
            from oracle import window as OW
            ocfg = O.GenCfg(**kw)
            win = (0, min(5, th), 0, min(5, tw))
            zb4 = z_band.unsqueeze(0)
            mb4 = None if maps_band is None else [m[None, None] for m in maps_band]
            t0 = time.perf_counter()
            err, trusted = OW.window_error(sd, ocfg, zb4, mb4, plan.out, th if world == 1 else th + 1, tw, win)
            parity = {"max_abs": err, "tolerance": 2e-2 if args.precision != "fp32" else 1e-3, "ok": err <= (2e-2 if args.precision != "fp32" else 1e-3),
                      "against": "oracle/itg_oracle.py forward_merged (CPU fp32) on the cropped noise window", "window_patches": list(win),
                      "trusted_patches": list(trusted), "oracle_s": time.perf_counter() - t0}
            if seam is not None:
                parity["seam"] = seam
                parity["ok"] = parity["ok"] and seam["ok"]
        if with_cpu and world == 1:
            torch.set_num_threads(os.cpu_count() or 1)
            v, threads, sample, _ = cpu_reference_rate(kw, th, tw, sd, budget_s=15.0)
            cpu_base = {"value": v, "unit": "MP/s", "cores": threads, "kind": "port", "sample": sample}
    entry = {"name": name, "value": value, "unit": "MP/s", "steps": steps, "ms_per_step": ms_per_step, "timed_ms": dev_ms,
             "config": {"workload": desc + (f"; x{world} row bands of that size, one per GPU (bands.RowBandSampler), per-layer halo rows over "
                                            + ("NVLink P2P (itg_halo_exchange)" if sampler.p2p else "torch.distributed send/recv") if world > 1 else ""),
                        "weights": "random init (reference init scheme)", "l2": f"flushed between steps ({L2_FLUSH_BYTES >> 20} MiB write)",
                        "launch": "CUDA graph replay" if graphed else "eager launches", "wall_ms_per_step_incl_flush": t_wall / steps * 1e3},
             "clocks": clk, "gpu_launches": launches,
             "e2e": {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world, "steps": e2e_steps,
                     "mode": "streaming public API (utils.generate_textures; N > 1: bands.RowBandSampler.generate): per step H2D of the noise from pinned memory + D2H of the "
                             "fp32 image into pinned memory, image k's D2H overlapped with pass k+1; wall clock over the steps incl. the last copy, max over ranks",
                     "sync_value": e2e_sync, "sync_mode": "one blocking sample_from_gen_PatchByPatch_test call + D2H per step (N = 1)",
                     "u8_value": e2e_u8, "u8_mode": "as value, with test_sample.py's 8-bit output stage on the device (d2h = 1/4 of d2h_bytes_per_step)"},
             "roofline": roof, "parity": parity}
    if cpu_base is not None:
        entry["cpu_baseline"] = cpu_base
    if sampler is not None:
        sampler.close()
    eng.drop_plans()
    del net, eng, plan
    torch.cuda.empty_cache()
    return entry


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="", choices=[""] + sorted(WORKLOADS),
                    help="default: the headline workload (" + HEADLINE + ") plus the other single-GPU BASELINE configs in `extra`")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"], help="N > 1: halo rows over peer-mapped memory (NVLink P2P) or NCCL send/recv")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="only the headline workload")
    ap.add_argument("--profile-out", default="")
    args = ap.parse_args()
    headline = args.workload or HEADLINE
    if args.impl == "reference":
        kw, th, tw, desc = WORKLOADS[headline]
        return run_reference(args, kw, th, tw, desc)
    args.warmup = max(args.warmup, 3)

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
        # pinned host buffers should live on the GPU's own NUMA node
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
    ctx = {"rank": rank, "world": world, "local": local, "dev": dev, "dist": dist,
           "flush": torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)}

    main_entry = run_workload(headline, args, ctx, args.steps, args.warmup, with_cpu=not args.no_cpu_baseline, with_parity=True,
                              profile_out=args.profile_out)
    extras = []
    if not args.workload and not args.no_extra:
        for name in EXTRA:
            # enough steps for >= 50 ms of timed device work per workload
            est = {"cfg2": 0.6, "cfg5band": 6.0}.get(name, 5.0)
            n = max(args.steps, int(60.0 / est) + 1)
            extras.append(run_workload(name, args, ctx, n, args.warmup, with_cpu=False, with_parity=True))

    if rank == 0:
        e = main_entry
        line = {"metric": "output megapixels/sec (Generator, local padding)", "value": e["value"], "unit": "MP/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": e["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": {"fp16": "f16", "fp32": "f32"}[args.precision],
                "data": "synthetic", "config": e["config"], "clocks": e["clocks"], "gpu_launches": e["gpu_launches"] + sum(x["gpu_launches"] for x in extras),
                "e2e": e["e2e"], "roofline": e["roofline"], "parity": e["parity"]}
        if "cpu_baseline" in e:
            line["cpu_baseline"] = e["cpu_baseline"]
        if extras:
            line["extra"] = extras
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
